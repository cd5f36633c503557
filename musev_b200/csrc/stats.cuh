// Launch accounting shared by all launchers: a process-wide launch counter (bench.py reports it as `gpu_launches`)
// and an optional per-category CUDA-event profile recorded on the launching stream (bench.py's roofline numbers).
#pragma once
#include <cuda_runtime.h>

namespace mvb {

enum KernelCategory { KC_GEMM = 0, KC_ATTENTION = 1, KC_TEMPORAL_ATTN = 2, KC_GROUPNORM = 3, KC_LAYERNORM = 4, KC_OTHER = 5, KC_COUNT = 6 };

void stats_note_launch(int category, int n = 1);
bool stats_profiling();
// RAII: brackets the launches issued inside its scope with two events when profiling is on.
struct ProfScope {
  cudaStream_t s;
  int cat;
  cudaEvent_t e0 = nullptr;
  ProfScope(cudaStream_t stream, int category, int launches = 1);
  ~ProfScope();
};

}  // namespace mvb
