// Host VM: program assembler, Rescue sponge, trace builder, LWE client ops. See vm.h for the
// reference lines each part mirrors.
#include "vm.h"
#include "../../../include/ezkvm_rescue_constants.h"
#include <cmath>
#include <sstream>

namespace ezk {

namespace {

struct Pair64 {
    uint64_t lo, hi;
};
const Pair64 kMds[16] = {EZK_RESCUE_MDS_INIT};
const Pair64 kArk[128] = {EZK_RESCUE_ARK_INIT};
inline Fp cst(const Pair64& p) { return Fp(((u128)p.hi << 64) | p.lo); }
const u128 kInvAlpha = ((u128)EZK_RESCUE_INV_ALPHA_HI << 64) | EZK_RESCUE_INV_ALPHA_LO;

constexpr size_t kCycle = 16, kRounds = 14;            // crypto/src/rescue.rs:12-14
constexpr size_t kPushAlign = 8;                       // vm/src/program/mod.rs:20
constexpr size_t kMinTrace = 16, kMaxDepth = 16;       // vm/src/processor/mod.rs:35-36

const char* op_name(uint8_t code) {
    switch (code) {
        case OP_NOOP: return "noop";
        case OP_PUSH: return "push";
        case OP_READ: return "read";
        case OP_READ2: return "read2";
        case OP_ADD: return "add";
        case OP_MUL: return "mul";
        case OP_SADD: return "sadd";
        case OP_SMUL: return "smul";
        case OP_ADD2: return "add2";
    }
    return "?";
}

// Inverse S-box x -> x^INV_ALPHA on the four state elements at once (crypto/src/rescue.rs:146-150 with
// INV_ALPHA :199).  The exponent is (2(M-1)+1)/3 = 0xAAAAAAAAAAAAAAAAAAAA8CAAAAAAAAAB: runs of the bit pair
// "10" around one irregular byte.  With A_k = x^("10" repeated k times), A_(j+k) = A_j^(4^k) * A_k, the chain
// below needs 127 squarings + 14 products instead of 127 + 63 for square-and-multiply, and every step is
// applied to the four independent lanes back to back so that their multiplications overlap in the core.
// The trace builder spends its time here: one call per executed operation, each depending on the last.
//
// fmul: the field product of f128_host.h (same folds, same result) as one x86-64 block: 6 `mulq` and ~40
// single-cycle instructions, where the compiler's code for the portable form is ~110 (measured: 8.9 ns against
// 15.7 ns per product with four lanes in flight).  Other hosts use the portable operator.
#if defined(__x86_64__) && defined(__GNUC__)
inline Fp fmul(Fp a, Fp b) {
    typedef unsigned long long ull;
    const ull a0 = (ull)a.v, a1 = (ull)(a.v >> 64), b0 = (ull)b.v, b1 = (ull)(b.v >> 64);
    ull w0, w1, w2, w3, t, u;
    unsigned char cf;
    __asm__(
        // 256-bit product w3:w2:w1:w0
        "movq %[a0], %%rax\n\t mulq %[b0]\n\t movq %%rax, %[w0]\n\t movq %%rdx, %[w1]\n\t"
        "movq %[a1], %%rax\n\t mulq %[b1]\n\t movq %%rax, %[w2]\n\t movq %%rdx, %[w3]\n\t"
        "movq %[a0], %%rax\n\t mulq %[b1]\n\t addq %%rax, %[w1]\n\t adcq %%rdx, %[w2]\n\t adcq $0, %[w3]\n\t"
        "movq %[a1], %%rax\n\t mulq %[b0]\n\t addq %%rax, %[w1]\n\t adcq %%rdx, %[w2]\n\t adcq $0, %[w3]\n\t"
        // H*45 = rdx:rax:u, shifted left by 40
        "movq %[w2], %%rax\n\t mulq %[k]\n\t movq %%rax, %[u]\n\t movq %%rdx, %[t]\n\t"
        "movq %[w3], %%rax\n\t mulq %[k]\n\t addq %[t], %%rax\n\t adcq $0, %%rdx\n\t"
        "shldq $40, %%rax, %%rdx\n\t"
        "shldq $40, %[u], %%rax\n\t"
        "shlq $40, %[u]\n\t"
        // L + (H*45 << 40) - H  ->  rdx:w1:w0, rdx < 2^47
        "addq %[u], %[w0]\n\t adcq %%rax, %[w1]\n\t adcq $0, %%rdx\n\t"
        "subq %[w2], %[w0]\n\t sbbq %[w3], %[w1]\n\t sbbq $0, %%rdx\n\t"
        // second fold: + rdx*(45*2^40 - 1)
        "leaq (%%rdx,%%rdx,4), %%rax\n\t leaq (%%rax,%%rax,8), %%rax\n\t"
        "movq %%rax, %[t]\n\t shlq $40, %%rax\n\t shrq $24, %[t]\n\t subq %%rdx, %%rax\n\t sbbq $0, %[t]\n\t"
        "addq %%rax, %[w0]\n\t adcq %[t], %[w1]\n\t setc %[cf]"
        : [w0] "=&r"(w0), [w1] "=&r"(w1), [w2] "=&r"(w2), [w3] "=&r"(w3), [t] "=&r"(t), [u] "=&r"(u), [cf] "=q"(cf)
        : [a0] "r"(a0), [a1] "r"(a1), [b0] "r"(b0), [b1] "r"(b1), [k] "r"(45ULL)
        : "rax", "rdx", "cc");
    u128 z = ((u128)w1 << 64) | w0;
    if (__builtin_expect(cf, 0)) z += (((u128)45) << 40) - 1;
    return Fp::reduce(z);
}
#else
inline Fp fmul(Fp a, Fp b) { return a * b; }
#endif

struct Lanes {
    Fp v[4];
};
inline Lanes sqr_n(Lanes a, int n) {
    for (int k = 0; k < n; k++)
        for (int i = 0; i < 4; i++) a.v[i] = fmul(a.v[i], a.v[i]);
    return a;
}
inline Lanes mul(Lanes a, const Lanes& b) {
    for (int i = 0; i < 4; i++) a.v[i] = fmul(a.v[i], b.v[i]);
    return a;
}
void inv_sbox(Fp s[4]) {
    Lanes x;
    for (int i = 0; i < 4; i++) x.v[i] = s[i];
    const Lanes a1 = sqr_n(x, 1);
    const Lanes a2 = mul(sqr_n(a1, 2), a1);
    const Lanes a4 = mul(sqr_n(a2, 4), a2);
    const Lanes a8 = mul(sqr_n(a4, 8), a4);
    const Lanes a16 = mul(sqr_n(a8, 16), a8);
    const Lanes a18 = mul(sqr_n(a16, 4), a2);
    const Lanes a36 = mul(sqr_n(a18, 36), a18);
    Lanes r = mul(sqr_n(a36, 8), a4);  // bits 127..48: "10" x 40
    r = mul(sqr_n(r, 1), x);           // 0x8C = 1000 1100
    r = mul(sqr_n(r, 4), x);
    r = mul(sqr_n(r, 1), x);
    r = sqr_n(r, 2);
    r = mul(sqr_n(r, 36), a18);        // bits 39..4: "10" x 18
    r = mul(sqr_n(r, 1), x);           // 0xB = 1011
    r = mul(sqr_n(r, 2), x);
    r = mul(sqr_n(r, 1), x);
    for (int i = 0; i < 4; i++) s[i] = r.v[i];
}

// The chain's exponent is written out by hand above; tie it to the constant of the reference once per process.
bool inv_sbox_matches_inv_alpha() {
    Fp probe[4] = {Fp(2), Fp(Fp::modulus() - 1), Fp((((u128)0x0123456789ABCDEFULL) << 64) | 0xFEDCBA9876543210ULL), Fp(0)};
    Fp want[4];
    for (int i = 0; i < 4; i++) want[i] = pow(probe[i], kInvAlpha);
    inv_sbox(probe);
    for (int i = 0; i < 4; i++)
        if (probe[i] != want[i] || fmul(fmul(probe[i], probe[i]), probe[i]) != pow(want[i], 3) ||
            fmul(want[i], want[(i + 1) & 3]) != want[i] * want[(i + 1) & 3])
            return false;
    return true;
}

void mds_mul(Fp s[4]) {  // crypto/src/rescue.rs:162-176
    Fp r[4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r[i] = r[i] + fmul(cst(kMds[i * 4 + j]), s[j]);
    for (int i = 0; i < 4; i++) s[i] = r[i];
}

std::vector<std::string> split_dots(const std::string& s) {
    std::vector<std::string> parts;
    size_t start = 0;
    for (;;) {
        size_t p = s.find('.', start);
        if (p == std::string::npos) {
            parts.push_back(s.substr(start));
            break;
        }
        parts.push_back(s.substr(start, p - start));
        start = p + 1;
    }
    return parts;
}

std::string join_dots(const std::vector<std::string>& p) {
    std::string s = p[0];
    for (size_t i = 1; i < p.size(); i++) s += "." + p[i];
    return s;
}

std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

[[noreturn]] void program_error(size_t step, const std::string& msg) {
    throw VmError{"program error at " + std::to_string(step) + ": " + msg};
}

// vm/src/program/parsers.rs + mod.rs:107-122
Operation parse_op(size_t step, const std::string& line) {
    auto op = split_dots(line);
    const std::string& name = op[0];
    if (name == "push") {
        if (op.size() == 1) program_error(step, "malformed instruction " + name + ", parameter is missing");
        if (op.size() > 2) program_error(step, "malformed instruction " + name + ", too many parameters provided");
        const std::string& p = op[1];
        size_t i = (!p.empty() && p[0] == '+') ? 1 : 0;  // Rust's u8::from_str accepts a leading '+'
        bool ok = i < p.size();
        unsigned v = 0;
        for (; ok && i < p.size(); i++) {
            if (p[i] < '0' || p[i] > '9') ok = false;
            else {
                v = v * 10 + (unsigned)(p[i] - '0');
                if (v > 255) ok = false;
            }
        }
        if (!ok) program_error(step, "malformed instruction " + name + ", parameter '" + p + "' is invalid");
        return Operation{OP_PUSH, (uint8_t)v};
    }
    uint8_t code;
    if (name == "read") code = OP_READ;
    else if (name == "read2") code = OP_READ2;
    else if (name == "add") code = OP_ADD;
    else if (name == "mul") code = OP_MUL;
    else if (name == "sadd") code = OP_SADD;
    else if (name == "smul") code = OP_SMUL;
    else if (name == "add2") code = OP_ADD2;
    else program_error(step, "instruction " + join_dots(op) + " is invalid");
    if (op.size() > 1) program_error(step, "malformed instruction " + name + ", too many parameters provided");
    return Operation{code, 0};
}

inline size_t pad_to_cycle(size_t len) { return len + (kCycle - (len % kCycle)); }  // program/mod.rs:124-126

}  // namespace

void host_field_products(Fp a, Fp b, Fp& portable, Fp& sponge, Fp& squared, Fp& inv_alpha_power) {
    portable = a * b;
    sponge = fmul(a, b);
    squared = square(a);
    Fp lanes[4] = {a, b, a + b, a - b};
    inv_sbox(lanes);
    inv_alpha_power = lanes[0];
}

std::string Operation::to_string() const {
    std::string s = op_name(code);
    if (code == OP_PUSH) s += "(" + std::to_string(value) + ")";
    return s;
}

void RescueSponge::update(uint8_t op_code, uint8_t op_value) {
    static const bool chain_ok = inv_sbox_matches_inv_alpha();
    if (!chain_ok) throw VmError{"internal error: inverse S-box chain does not compute x^INV_ALPHA"};
    if (step % kCycle < kRounds) {  // rescue.rs:42-54, 102-118
        const Pair64* ark = &kArk[(step % kCycle) * 8];
        for (int i = 0; i < 4; i++) state[i] = fmul(fmul(state[i], state[i]), state[i]);
        mds_mul(state);
        for (int i = 0; i < 4; i++) state[i] = state[i] + cst(ark[i]);
        state[0] = state[0] + Fp::from_u64(op_code);
        state[1] = state[1] + Fp::from_u64(op_value);
        inv_sbox(state);
        mds_mul(state);
        for (int i = 0; i < 4; i++) state[i] = state[i] + cst(ark[4 + i]);
    } else {
        state[2] = Fp();
        state[3] = Fp();
    }
    step++;
}

Program Program::compile(const std::string& source) {
    std::vector<std::string> tokens;
    std::istringstream in(source);
    std::string raw;
    while (std::getline(in, raw)) {
        std::string line = trim(raw);
        if (line.empty() || line[0] == '#') continue;
        size_t pos = line.find('#');
        std::string code_line = trim(pos == std::string::npos ? line : line.substr(0, pos));
        if (!code_line.empty()) tokens.push_back(code_line);
    }
    if (tokens.empty()) program_error(0, "a program must contain at least one instruction");

    Program p;
    const Operation noop{OP_NOOP, 0};
    for (size_t i = 0; i < tokens.size(); i++) {
        Operation op = parse_op(i + 1, tokens[i]);
        if (op.code == OP_PUSH) {  // mod.rs:68-72
            size_t pad = (kPushAlign - p.code.size() % kPushAlign) % kPushAlign;
            p.code.resize(p.code.size() + pad, noop);
        }
        if (p.code.size() % kCycle >= kRounds) p.code.resize(pad_to_cycle(p.code.size()), noop);  // mod.rs:76-79
        p.code.push_back(op);
    }
    p.code.resize(pad_to_cycle(p.code.size()), noop);  // mod.rs:85-86 (pads a full cycle when already aligned)
    RescueSponge sponge;
    p.sponge_states.resize(4 * p.code.size());
    for (size_t i = 0; i < p.code.size(); i++) {
        sponge.update(p.code[i].code, p.code[i].value);
        for (int j = 0; j < 4; j++) p.sponge_states[4 * i + j] = sponge.state[j];
    }
    p.hash[0] = sponge.state[0];
    p.hash[1] = sponge.state[1];
    return p;
}

std::string Program::to_string() const {
    std::string s;
    for (size_t i = 0; i < code.size(); i++) {
        if (i) s += " ";
        s += code[i].to_string();
    }
    return s;
}

ExecutionTrace execute(const Program& program, const std::vector<uint8_t>& pub, const std::vector<Fp>& secret,
                       const LweParams& lwe, uint64_t last_row_seed) {
    const size_t lw = lwe.lwe_size();
    const size_t num_ops = program.code.size();
    // rows 0..num_ops are produced by execution; capacity doubling rule of chiplets.rs:80-90
    size_t capacity = kMinTrace;
    std::vector<std::vector<Fp>> regs(kMaxDepth, std::vector<Fp>(num_ops + 1));
    std::vector<Fp> depth_col(num_ops + 1), flag_col(num_ops + 1);
    std::vector<std::vector<Fp>> bits(5, std::vector<Fp>(num_ops + 1)), sponge_cols(4, std::vector<Fp>(num_ops + 1));
    RescueSponge sponge;
    size_t depth = 0, tape_a = 0, tape_b = 0, clk = 0;
    const size_t num_ct = lw ? secret.size() / lw : 0;
    const bool have_states = program.sponge_states.size() == 4 * num_ops;

    auto stack_err = [&](const Operation& op, const std::string& what) {
        throw VmError{"stack error at " + std::to_string(clk) + ": " + what};
        (void)op;
    };
    auto shift_left = [&](const Operation& op, size_t start, size_t pc) {  // stack.rs:220-238
        if (depth < pc) stack_err(op, op.to_string() + " operation stack underflow");
        for (size_t i = start; i < depth; i++) regs[i - pc][clk] = regs[i][clk - 1];
        for (size_t i = depth - pc; i < depth; i++) regs[i][clk] = Fp();
        depth -= pc;
    };
    auto shift_right = [&](const Operation& op, size_t pc) {  // stack.rs:240-255
        depth += pc;
        if (depth > kMaxDepth) stack_err(op, op.to_string() + " operation stack overflow");
        for (size_t i = 0; i < depth - pc; i++) regs[i + pc][clk] = regs[i][clk - 1];
    };

    for (const Operation& op : program.code) {
        clk++;                                  // system.advance_step / stack.advance_clock
        if (clk >= capacity) capacity *= 2;     // ensure_trace_capacity
        switch (op.code) {                      // stack.rs:48-70
            case OP_NOOP:
                for (size_t i = 0; i < depth; i++) regs[i][clk] = regs[i][clk - 1];
                break;
            case OP_PUSH:
                shift_right(op, 1);
                regs[0][clk] = Fp::from_u64(op.value);
                break;
            case OP_READ:
                shift_right(op, 1);
                if (tape_a >= pub.size()) stack_err(op, "no more inputs to " + op.to_string());
                regs[0][clk] = Fp::from_u64(pub[tape_a++]);
                break;
            case OP_READ2: {
                if (tape_b >= num_ct) stack_err(op, "no more inputs to " + op.to_string());
                const Fp* ct = &secret[tape_b * lw];
                tape_b++;
                shift_right(op, lw);
                for (size_t i = 0; i < lw; i++) regs[i][clk] = ct[i];
                break;
            }
            case OP_ADD:
            case OP_MUL: {
                if (depth < 2) stack_err(op, op.to_string() + " operation stack underflow");
                Fp x = regs[0][clk - 1], y = regs[1][clk - 1];
                regs[0][clk] = op.code == OP_ADD ? x + y : x * y;
                shift_left(op, 2, 1);
                break;
            }
            case OP_SADD:
            case OP_SMUL: {
                if (depth < lw + 1) stack_err(op, op.to_string() + " operation stack underflow");
                Fp scalar = regs[0][clk - 1];
                for (size_t i = 0; i < lw; i++) {
                    Fp ct = regs[1 + i][clk - 1];
                    if (op.code == OP_SMUL) regs[i][clk] = ct * scalar;                       // server_key.rs:116-124
                    else regs[i][clk] = (i == lwe.k) ? ct + Fp::from_u64(lwe.delta) * scalar : ct;  // :78-83,104-114
                }
                shift_left(op, lw + 1, 1);
                break;
            }
            case OP_ADD2: {
                if (depth < lw * 2) stack_err(op, op.to_string() + " operation stack underflow");
                for (size_t i = 0; i < lw; i++) regs[i][clk] = regs[i][clk - 1] + regs[i + lw][clk - 1];  // :89-102
                shift_left(op, lw * 2, lw);
                break;
            }
            default:
                throw VmError{"unknown opcode"};
        }
        depth_col[clk] = Fp::from_u64(depth);                                        // stack.rs:278-280
        for (int i = 0; i < 5; i++) bits[i][clk - 1] = Fp::from_u64((op.code >> i) & 1);  // decoder.rs:68-76
        if (sponge.step % kCycle >= kRounds && op.code != OP_NOOP)                   // chiplets.rs:92-95
            throw VmError{"chiplets error at " + std::to_string(clk) + ": expected noop but was " + op.to_string()};
        if (have_states) {  // the chain compile() already ran over the same operations
            for (int i = 0; i < 4; i++) sponge.state[i] = program.sponge_states[4 * (clk - 1) + i];
            sponge.step++;
        } else {
            sponge.update(op.code, op.value);
        }
        flag_col[clk - 1] = Fp(1);                                                   // chiplets.rs:99-105
        for (int i = 0; i < 4; i++) sponge_cols[i][clk] = sponge.state[i];           // chiplets.rs:107-109
    }
    if (clk % kCycle != 0)  // chiplets.rs:41-43
        throw VmError{"chiplets error at " + std::to_string(clk) + ": trace length should be a multiple of " +
                      std::to_string(kCycle) + ", but was " + std::to_string(clk)};

    ExecutionTrace t;
    for (size_t i = 0; i < kMaxDepth; i++) t.outputs[i] = regs[i][clk];
    size_t n = 1;
    while (n < capacity + 1) n <<= 1;  // mod.rs:74 (NUM_RAND_ROWS = 1)
    t.n = n;
    t.columns.assign(28, std::vector<Fp>(n));
    for (size_t i = 0; i < n; i++) t.columns[0][i] = Fp::from_u64(i);  // system.rs:19-28
    auto fill = [&](std::vector<Fp>& dst, const std::vector<Fp>& src, bool repeat_last) {
        for (size_t i = 0; i <= clk; i++) dst[i] = src[i];
        for (size_t i = clk + 1; i < n; i++) dst[i] = repeat_last ? src[clk] : Fp();
    };
    for (int i = 0; i < 5; i++) fill(t.columns[1 + i], bits[i], false);          // decoder.rs:31-47
    fill(t.columns[6], flag_col, false);                                         // chiplets.rs:50-53
    for (int i = 0; i < 4; i++) fill(t.columns[7 + i], sponge_cols[i], true);    // chiplets.rs:45-48
    fill(t.columns[11], depth_col, true);                                        // stack.rs:88-91
    for (size_t i = 0; i < kMaxDepth; i++) fill(t.columns[12 + i], regs[i], true);  // stack.rs:83-86
    SplitMix64 rng(last_row_seed);
    for (auto& col : t.columns) col[n - 1] = rng.next_fp_nonzero();  // mod.rs:86-92
    return t;
}

// ---------------------------------------------------------------------------------------------
namespace {
struct ProgramBuilder {
    std::string src;
    size_t len = 0;  // compiled length so far, following Program::compile's padding rules
    void emit(const char* tok, bool is_push = false) {
        if (is_push) len += (kPushAlign - len % kPushAlign) % kPushAlign;
        if (len % kCycle >= kRounds) len = pad_to_cycle(len);
        len++;
        src += tok;
        src += '\n';
    }
};
}  // namespace

SyntheticCase make_synthetic(int kind, unsigned log_n, const LweParams& lwe, uint64_t seed) {
    // n = 4 * 2^floor(log2 Lp)  =>  need 2^(log_n-2) <= Lp < 2^(log_n-1); aim at ~7/8 of the upper bound
    const size_t hi = (size_t)1 << (log_n - 1), lo = (size_t)1 << (log_n - 2);
    const size_t target = lo + (hi - lo) * 3 / 4;
    SplitMix64 rng(seed);
    ProgramBuilder b;
    size_t reads = 0, read2s = 0;
    auto scalar_block = [&](bool first) {
        if (first) {
            b.emit("read");
            reads++;
        }
        b.emit("read");
        reads++;
        b.emit((rng.next() & 1) ? "add" : "mul");
    };
    auto ct_block = [&](bool first) {
        b.emit("read2"), read2s++;
        b.emit("read"), reads++;
        b.emit("smul");
        if (!first) b.emit("add2");
    };
    size_t iter = 0;
    if (kind == 1) {
        scalar_block(true);
        while (b.len + 40 < target) {
            if (++iter % 32 == 0) {
                std::string t = "push." + std::to_string(rng.next() & 0xFF);
                b.emit(t.c_str(), true);
                b.emit((rng.next() & 1) ? "add" : "mul");
            } else {
                scalar_block(false);
            }
        }
    } else if (kind == 2) {
        ct_block(true);
        while (b.len + 40 < target) {
            if (++iter % 16 == 0) {
                b.emit("read"), reads++;
                b.emit("sadd");
            } else {
                ct_block(false);
            }
        }
    } else {
        ct_block(true);
        while (b.len + 40 < target) {
            ++iter;
            uint64_t r = rng.next() % 4;
            if (r == 0) {
                ct_block(false);
            } else {
                // scalar work on top of the ciphertext, then fold the scalar into it
                b.emit("read"), reads++;
                b.emit("read"), reads++;
                b.emit((rng.next() & 1) ? "add" : "mul");
                if (iter % 8 == 0) {
                    std::string t = "push." + std::to_string(rng.next() & 0xFF);
                    b.emit(t.c_str(), true);
                    b.emit("mul");
                }
                b.emit((rng.next() & 1) ? "smul" : "sadd");
            }
        }
    }
    SyntheticCase c;
    c.program = Program::compile(b.src);
    c.pub.resize(reads);
    for (auto& v : c.pub) v = (uint8_t)(rng.next() & 0xFF);
    c.secret.resize(read2s * lwe.lwe_size());
    for (auto& v : c.secret) v = rng.next_fp();
    return c;
}

// ---------------------------------------------------------------------------------------------
LweKey lwe_keygen(const LweParams& p, uint64_t seed) {  // server_key.rs:19-28
    SplitMix64 rng(seed);
    LweKey k;
    k.params = p;
    for (uint32_t i = 0; i < p.k; i++) k.key.push_back(Fp::from_u64(rng.next() & 1));
    return k;
}

std::vector<Fp> lwe_encrypt(const LweKey& k, uint8_t value, uint64_t seed) {  // server_key.rs:41-62
    SplitMix64 rng(seed);
    std::vector<Fp> ct;
    for (uint32_t i = 0; i < k.params.k; i++) ct.push_back(rng.next_fp());
    // Box-Muller normal sample (the reference uses rand_distr::Normal with thread_rng)
    double u1 = ((rng.next() >> 11) + 1.0) / 9007199254740993.0, u2 = (rng.next() >> 11) / 9007199254740992.0;
    double noise = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2) * k.params.std_dev;
    Fp scaled = Fp::from_u64((uint64_t)std::llround(std::fabs(noise)));
    Fp body;
    for (uint32_t i = 0; i < k.params.k; i++) body = body + ct[i] * k.key[i];
    body = body + Fp::from_u64(k.params.delta) * Fp::from_u64(value);
    body = noise > 0.0 ? body + scaled : body - scaled;
    ct.push_back(body);
    return ct;
}

uint8_t lwe_decrypt(const LweKey& k, const Fp* ct) {  // server_key.rs:64-76
    Fp mask;
    for (uint32_t i = 0; i < k.params.k; i++) mask = mask + ct[i] * k.key[i];
    Fp m = ct[k.params.k] - mask;
    u128 log2_delta = (u128)std::log2((double)k.params.delta);
    u128 round_bit = (m.v >> (log2_delta - 1)) & 1;
    return (uint8_t)((m.v >> log2_delta) + round_bit);
}

}  // namespace ezk
