// nm_registry.h -- internal interface between the C-ABI translation unit (nm_api.cu) and the
// per-shape-class translation units (nm_spec_tu.cu compiled once per generated spec header).
#pragma once
#include <cuda_runtime.h>
#include "nm_rollout.h"

struct NmFwdCall {
  const float* params; const float* x0; const float* const* exo;
  float* x_traj; float* u_traj; float* y_traj; float* checkpoints;
  void* workspace; size_t workspace_bytes; int64_t B; cudaStream_t stream;
  int fused_terms;                                 // in: plan terms match the entry's compiled term structure
  double* term_partials; int grid;                 // out: per-CTA term sums (fused), CTAs launched
  int extra_launches;                              // out: kernels launched besides weight prep + rollout
  int scored;                                      // out: the rollout kernel wrote term_partials (else the API scores the terms)
};
struct NmBwdCall {
  const float* params; const float* const* exo;
  const float* x_traj; const float* u_traj; const float* y_traj; const float* checkpoints;
  const float* grad_scalars; const float* gx_ext; const float* gu_ext; const float* gy_ext;
  float* grad_x0;
  const float* G_x; const float* G_u; const float* G_y;   // precomputed term gradients or null
  int fused_terms;
  void* workspace; size_t workspace_bytes; int64_t B; int64_t b_total; cudaStream_t stream;
  float* grad_partials; int grid;                  // out: per-CTA gradient blocks, CTAs launched
  int extra_launches;                              // out: kernels launched besides weight prep + adjoint
};

struct NmKernelEntry {
  const char* signature;
  const char* term_signature;                     // compiled term structure ("" = none)
  int (*workspace)(const NmPlan*, int64_t B, size_t* fwd_bytes, size_t* bwd_bytes);
  int (*forward)(const NmPlan*, NmFwdCall*);      // prep + rollout kernel; returns NM_OK / NM_ERR_*
  int (*backward)(const NmPlan*, NmBwdCall*);     // prep + adjoint kernel
  const char* info;                               // tile / threads / smem, for diagnostics
  int chk_floats_per_step;                        // checkpoint floats per trajectory-step the kernels of this entry keep
                                                  // (-1: generic rule -- stage derivatives for MLP right-hand sides)
  NmKernelEntry* next;
};

extern "C" void nm_register_kernel(NmKernelEntry* e);
extern "C" void nm_set_error(const char* msg);
