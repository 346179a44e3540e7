// Internal kernel-launch interfaces shared between the translation units of libpacoh_b200.
#pragma once
#include "common.cuh"

namespace pacoh {

// ---- MLP forward / backward (mlp.cu, mlp_generic.cu) -------------------------------------------------
struct MlpArgs {
  const float* theta;     // (P, D)
  const float* x;         // (T_total, n, d)
  const int* task_idx;    // (T)
  int P, T, n, d, D;
  NetDev net[2];          // nets handled by this launch (blockIdx.z)
  float* out[2];          // fwd: (P, Q, out_dim)
  const float* dout[2];   // bwd: (P, Q, out_dim)
  float* partial[2];      // bwd: (chunks, P, net.total)
};
int launch_mlp_fast(const MlpArgs& a, int nets, int chunks, bool bwd, cudaStream_t st);
int launch_mlp_tc_fwd(const MlpArgs& a, int nets, int chunks, cudaStream_t st);   // tcgen05 forward (mlp_tc.cu)
int launch_mlp_tc_bwd(const MlpArgs& a, int nets, int chunks, cudaStream_t st);   // tcgen05 + CUDA-core backward (mlp_tc_bwd.cu)
int mlp_tc_bwd_slots(int P, int nets, int Q, int* grid_out, int* per_cta_out);    // partial slots of its persistent schedule
int launch_mlp_generic(const MlpArgs& a, int net_index, int chunks, bool bwd, float* scratch, cudaStream_t st);
size_t mlp_generic_scratch_floats(const NetDev& net, int P, int Q);

// ---- batched GP marginal log-likelihood + analytic gradient (gp_mll.cu) --------------------------------
struct GpArgs {
  const float* theta;     // (P, D): lengthscale_raw, noise_raw, outputscale_raw, constant_mean are read from here
  const float* x;         // raw inputs (covar SE uses them as features)
  const float* y;         // (T_total, n)
  const int* task_idx;    // (T)
  const int* task_n;      // (T_total) points per task for ragged batches (rows n_t .. n-1 are padding), or nullptr
  const float* mean;      // (P, Q) or nullptr (zero / constant mean)
  const float* feat;      // (P, Q, F) or nullptr (SE on raw inputs)
  float* dmean;           // (P, Q) or nullptr
  float* dfeat;           // (P, Q, F) or nullptr
  float* mll;             // (P, T)
  float* dhyp;            // (P, T, HYP): [dl_raw(F), dnoise_raw, doscale_raw, dconst_mean]
  int* info;              // (P, T) or nullptr
  int P, T, n, d, F, D;
  int mean_kind, covar_kind, has_oscale;
  float noise_floor;
  int off_ls, off_noise, off_oscale, off_const_mean;
  int values_only;        // 1: only mll / info are needed (predictive path): the large-n path then stops after the factorisation
};
constexpr int kMaxGpN = 128;   // register kernel: n <= 64; tensor-memory kernel: 32 < n <= 128 (feature dim <= 4)
constexpr int kMaxBigN = 4096; // blocked tensor-core Cholesky path (gp_big.cu): 64 < n <= 4096, feature dim <= 4
constexpr int kMaxGpF = 16;    // feature dim
__host__ __device__ inline int gp_hyp_stride(int F) { return F + 3; }
int launch_gp_mll(const GpArgs& a, cudaStream_t st);
int launch_gp_mll_tc(const GpArgs& a, cudaStream_t st);   // gp_tc.cu: tcgen05 / tensor-memory version, 32 < n <= 64, F <= 4
// gp_big.cu: blocked Cholesky / inverse on the tensor cores for n > 64 (BASELINE config #5); scratch from the caller
size_t gp_big_workspace_bytes(int n, long long matrices);
int launch_gp_mll_big(const GpArgs& a, void* ws, size_t ws_bytes, cudaStream_t st);
void gp_big_layout_debug(int n, long long matrices, long long* out);
bool gp_use_big(int n, int F);   // capi.cu: which path a (n, F) problem takes

// ---- eval-mode posterior + evaluation metrics (gp_post.cu): C-ABI entry points only, see include/pacoh_b200.h
// ---- reductions / elementwise (finalize.cu) ------------------------------------------------------------
int launch_reduce_partials(const float* partial, int chunks, int P, int total, float* dtheta, int D, int dst_off,
                           cudaStream_t st);
int launch_reduce_hyp(const GpArgs& a, float* dtheta, float* mll_sum, cudaStream_t st);

}  // namespace pacoh
