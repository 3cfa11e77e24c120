// Host side of the ingest cast: fp32 -> fp16 (round to nearest even, what numpy's astype(float16) does,
// qdrant_indexer.py:423-441) on the CPU, and a small worker pool that lets an upload convert its rows into pinned staging
// memory with several threads while the previous chunk is on its way to the device. Plain C++ (compiled by g++, F16C where
// the CPU has it): no CUDA types here.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <functional>

namespace vrag {

// dst[i] = fp16(src[i]), i in [0, n); one thread.
void host_f32_to_f16(const float* src, uint16_t* dst, size_t n, bool force_scalar = false);

// Runs fn(0) ... fn(n_tasks - 1) on the pool's workers and the calling thread; returns when all are done.
// Calls from different threads are serialised.
void host_parallel_for(int n_tasks, const std::function<void(int)>& fn);
int host_pool_threads();   // workers + the caller (VRAG_HOST_THREADS, default min(8, hardware threads))

}  // namespace vrag
