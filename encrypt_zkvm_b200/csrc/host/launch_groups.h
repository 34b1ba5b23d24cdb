// How many trace columns of a HOST trace go into one interpolation + LDE launch (prover.cu, host-trace branch).
//
// Columns arrive one by one on the copy stream; their transforms run on the compute stream.  Large launches are 10-15 %
// more efficient than 2-column ones, but a launch cannot start before its last column is on the device.  The rule:
//   * nothing queued on the compute stream  -> launch what has arrived at once (the first column; an upload-bound host);
//   * one launch running, nothing behind it -> launch only when at least as many columns have arrived as that launch
//     holds (or when every column has been sent), so the groups grow while the upload runs ahead: 1, 1, 2, 2, 3, ...;
//   * two or more launches pending          -> keep collecting;
//   * `cap` columns have arrived            -> launch (the transform scratch is sized for `cap` columns).
// Host-only (no CUDA): ezk_selftest_launch_groups replays it against a small discrete-event model.
#pragma once
#include <cstdint>

namespace ezk {

// avail: columns that have arrived and are not launched yet; pending: launches queued or running on the compute stream;
// running_cols: columns of the oldest pending launch; all_sent: no further column will arrive
inline bool host_group_ready(uint32_t avail, uint32_t pending, uint32_t running_cols, uint32_t cap, bool all_sent) {
    if (avail == 0) return false;
    if (pending == 0 || avail >= cap) return true;
    if (pending == 1) return avail >= running_cols || all_sent;
    return false;
}

}  // namespace ezk
