// temp_b200 -- declarations shared by the translation units of libtemp_b200.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include "temp_b200.h"

namespace temp_internal {

// thread-local last-error string behind temp_last_error_string()
int fail(int code, const char* fmt, const char* a = "", long b = 0);
int cuda_fail(cudaError_t e, const char* what);

// tcgen05 path (tc_kernels.cu).  *_supported() say whether a launch meets the path's preconditions
// (d == 128, 1x1 relation blocks, one un-decayed dense term, packed operand images present ...);
// launches that do not are run by the fp32 SIMT kernels of temp_kernels.cu.
bool tc_layer_supported(const TempRgcnLayerArgs* a);
int tc_launch_layer(const TempRgcnLayerArgs* a, cudaStream_t st);
int tc_gather_grid(const TempRgcnLayerArgs* a);
int tc_launch_gather(const TempRgcnLayerArgs* a, cudaStream_t st);
int tc_gather_launches(const TempRgcnLayerArgs* a);
bool tc_scan_supported(const TempGruScanArgs* a);
int tc_launch_scan(const TempGruScanArgs* a, cudaStream_t st);
// gru_scan_tm_kernel (tc_scan2.cu): one recurrent cell, partitions of <= 48 rows per step; tc_launch_scan dispatches to it
bool tc_scan2_supported(const TempGruScanArgs* a);
int tc_launch_scan2(const TempGruScanArgs* a, cudaStream_t st);
int tc_pack_weights(const float* w_kn, int k, int n, void* packed, cudaStream_t st);
// 64-row tile kernel for the other widths (tc_wide.cu): d % 4 == 0, d <= 256, 1x1 / 2x2 / 4x4 relation blocks;
// tc_layer_supported / tc_launch_layer / tc_launch_gather / tc_pack_weights dispatch to these
bool tcw_layer_supported(const TempRgcnLayerArgs* a);
int tcw_launch_layer(const TempRgcnLayerArgs* a, cudaStream_t st);
int tcw_launch_gather(const TempRgcnLayerArgs* a, cudaStream_t st);
int64_t tcw_packed_bytes(int k, int n);
int tcw_pack_weights(const float* w_kn, int k, int n, void* packed, cudaStream_t st);
// gru_step_tcw_kernel: one GRU step per launch (64 rows x 32 hidden columns per CTA) for the widths the chain-partitioned
// 128-wide scans do not take; a scan of such steps is a chain of programmatic dependent launches
int64_t tcw_packed_gru_bytes(int d);
int tcw_pack_gru_weights(const float* whh_t, int d, void* packed, cudaStream_t st);
bool tcw_gru_supported(const TempGruArgs* g);
int tcw_launch_gru(const TempGruArgs* g, cudaStream_t st);
bool tcw_scan_supported(const TempGruScanArgs* a);
int tcw_launch_scan(const TempGruScanArgs* a, cudaStream_t st);
int tcw_scan_launches(const TempGruScanArgs* a);   // 1 (cooperative gru_scan_tcw_kernel, d <= 224) or one per non-empty step
int tc_pack_gru_weights(const float* whh_t, int d, void* packed, cudaStream_t st);

}  // namespace temp_internal
