// winattn_tc.cuh -- tcgen05/TMEM/TMA window attention (placeholder until the kernel lands).
#pragma once
#include "../../include/mmn_b200.h"
#include <cuda_runtime.h>
namespace mmn { namespace tc {
inline bool winattn_supported(const mmn_winattn_desc*) { return false; }
inline bool winattn_bwd_supported(const mmn_winattn_desc*) { return false; }
inline const char* why_not(const mmn_winattn_desc*) { return "tcgen05 path not built"; }
inline int winattn_fwd(const mmn_winattn_desc*, const void*, const void*, const void*, const float*, const float*, const float*,
                       void*, float*, cudaStream_t, char*, size_t) { return MMN_ERR_UNSUPPORTED; }
inline int winattn_bwd(const mmn_winattn_desc*, const void*, const void*, const void*, const float*, const float*, const float*,
                       const void*, const float*, const void*, void*, void*, void*, float*, float*, cudaStream_t, char*, size_t) { return MMN_ERR_UNSUPPORTED; }
}}
