// zero_fill.h -- zeroing of small accumulators / counters on a stream with a KERNEL instead of cudaMemsetAsync.
// Inside a CUDA graph a memset node sits between kernel nodes with ~2 us before it and ~5 us after it (kernel -> kernel
// edges cost ~0.1 us): tools/graph_timeline.py measured 8.5-9 us from the end of the kernel before the work-counter
// memset to the start of the attention kernel after it, twice per window-attention block and step.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace mmn {
// p: 4-byte aligned; words: number of 32-bit words to clear.
cudaError_t zero_words_async(void* p, size_t words, cudaStream_t st);
}  // namespace mmn
