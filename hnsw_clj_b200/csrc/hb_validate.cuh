// hb_validate.cuh — device-side validation of caller-supplied index arrays (hb_validate.cu).
#pragma once
#include "hb_common.cuh"

namespace hb {

// Each launcher ORs 1 into *flag (device int32, zeroed by the caller) when a value is outside [lo, hi).
void launch_check_range_i32(const int32_t *p, int64_t n, int64_t lo, int64_t hi, int32_t *flag);
void launch_check_range_i64(const int64_t *p, int64_t n, int64_t lo, int64_t hi, int32_t *flag);
// off[0] == 0, off non-decreasing over `count` entries, off[count-1] == total (total < 0: not checked)
void launch_check_offsets(const int64_t *off, int64_t count, int64_t total, int32_t *flag);
// CSR adjacency over n nodes: every id in [0, n); degree <= max_deg (0: not checked).  Offsets must have been checked.
void launch_check_csr(const int64_t *off, const int32_t *ids, int64_t n, int64_t max_deg, int32_t *flag);

}  // namespace hb
