// pg_scan.cuh -- device-wide exclusive scan (u32 counts -> u64 offsets), three small kernels.
#pragma once
#include "pg_common.cuh"
namespace pg {
size_t scan_workspace_bytes(uint64_t n);
// out[i] = sum(in[0..i)), *total (device pointer) = sum of all.  ws from scan_workspace_bytes.
int exclusive_scan_u32(const unsigned *in, unsigned long long *out, uint64_t n, unsigned long long *d_total,
                       void *ws, size_t wsBytes, cudaStream_t stream, uint64_t *launches);
}  // namespace pg
