// Run-time specialisation of the fused kernel through NVRTC (see jit.cpp).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "errors.hpp"
#include "kernels.cuh"

namespace chdb {

enum class JitMode { Never, Auto, Always };
// CHDB_JIT = 0 | never : interpreter kernel only;  unset | 1 | auto : specialise batches of at least
// kJitAutoRows rows;  always : specialise every launch (tests).
JitMode jit_mode();
constexpr int64_t kJitAutoRows = 1 << 18;

struct JitKernel {   // one NVRTC module: the stream kernel (single-batch and many-batch entry points) specialised for one program
  std::vector<char> cubin;
  cudaLibrary_t library = nullptr;
  cudaKernel_t stream = nullptr, stream_many = nullptr, select = nullptr, gather = nullptr;
};

bool jit_available(std::string* why);
std::string jit_prologue(const KernelParams& kp, bool has64, int min_blocks);
// Cached; returns nullptr (and the reason) when specialisation is impossible -> use the interpreter.
const JitKernel* jit_get(const KernelParams& kp, bool has64, int min_blocks, std::string* err);
// mode: StreamMode (0 fused, 1 select, 2 gather)
cudaError_t jit_launch_stream(const JitKernel* k, const KernelParams& p, const TilePlan& tp, int mode, unsigned grid, cudaStream_t stream);
void jit_stats(int64_t* compiles, double* seconds);
std::vector<char> jit_compile_offline(const KernelParams& kp, bool has64, std::string* log);

}  // namespace chdb
