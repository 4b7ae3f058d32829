// model.cu -- model upload: K_i = sf2*exp(-1/2 dist) + (sn2+eps_f32) I, blocked Cholesky K = L L^T,
// W = L^-1 (lower-triangular inverse) and alpha = K^-1 (Y_i - m0_i), all FP64 on the device.
// Replaces the host-side jnp.linalg.inv(Kopt) of models/GP_Safe.py:226-232 for the grid path.
#include "common.cuh"
#include <math.h>
#include <string.h>

#define NB 32   // factorisation block size

// K[g][r][c], padded to npad: identity on the padded diagonal so the padded matrix stays SPD.
__global__ void k_build_K(ModelSpec ms, double* __restrict__ K) {
  const int g = blockIdx.z;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ms.npad || c >= ms.npad) return;
  double v;
  if (r < ms.n && c < ms.n) {
    double s = 0.0;
    for (int k = 0; k < ms.d; ++k) {
      const double df = ms.Xn[r * ms.d + k] - ms.Xn[c * ms.d + k];
      s += df * df * ms.inv_ell[g][k];
    }
    v = ms.sf2[g] * exp(-0.5 * s);
    if (r == c) v += ms.sn2[g];
  } else {
    v = (r == c) ? 1.0 : 0.0;
  }
  K[((size_t)g * ms.npad + r) * ms.npad + c] = v;
}

// factor the NBxNB diagonal block kb in shared memory (right-looking, unblocked)
__global__ void __launch_bounds__(NB * NB) k_chol_diag(double* __restrict__ A, int npad, int kb, int* __restrict__ info) {
  __shared__ double a[NB][NB + 1];
  const int g = blockIdx.x;
  double* Ag = A + (size_t)g * npad * npad;
  const int tx = threadIdx.x, ty = threadIdx.y;   // ty = row, tx = col
  const int o = kb * NB;
  a[ty][tx] = Ag[(size_t)(o + ty) * npad + o + tx];
  __syncthreads();
  for (int j = 0; j < NB; ++j) {
    if (tx == j && ty == j) {
      const double p = a[j][j];
      if (!(p > 0.0)) { atomicExch(info + g, o + j + 1); a[j][j] = 1.0; }
      else a[j][j] = sqrt(p);
    }
    __syncthreads();
    if (tx == j && ty > j) a[ty][j] /= a[j][j];
    __syncthreads();
    if (tx > j && ty >= tx) a[ty][tx] -= a[ty][j] * a[tx][j];
    __syncthreads();
  }
  Ag[(size_t)(o + ty) * npad + o + tx] = (tx <= ty) ? a[ty][tx] : 0.0;
}

// panel below the diagonal block:  X * Lkk^T = A_panel   (one thread per row)
__global__ void __launch_bounds__(128) k_chol_panel(double* __restrict__ A, int npad, int kb) {
  __shared__ double l[NB][NB + 1];
  const int g = blockIdx.y;
  double* Ag = A + (size_t)g * npad * npad;
  const int o = kb * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) l[e / NB][e % NB] = Ag[(size_t)(o + e / NB) * npad + o + e % NB];
  __syncthreads();
  const int row = o + NB + blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= npad) return;
  double x[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) x[j] = Ag[(size_t)row * npad + o + j];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    double s = x[j];
#pragma unroll
    for (int t = 0; t < j; ++t) s -= x[t] * l[j][t];
    x[j] = s / l[j][j];
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) Ag[(size_t)row * npad + o + j] = x[j];
}

// trailing update  A[bi][bj] -= P_bi * P_bj^T  for lower tiles bi >= bj (NBxNB tiles)
__global__ void __launch_bounds__(NB * NB) k_chol_update(double* __restrict__ A, int npad, int kb) {
  const int bj = blockIdx.x, bi = blockIdx.y;
  if (bj > bi) return;
  __shared__ double pi[NB][NB + 1], pj[NB][NB + 1];
  const int g = blockIdx.z;
  double* Ag = A + (size_t)g * npad * npad;
  const int o = kb * NB;
  const int r0 = o + NB + bi * NB, c0 = o + NB + bj * NB;
  const int tx = threadIdx.x, ty = threadIdx.y;
  pi[ty][tx] = Ag[(size_t)(r0 + ty) * npad + o + tx];
  pj[ty][tx] = Ag[(size_t)(c0 + ty) * npad + o + tx];
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int t = 0; t < NB; ++t) s += pi[ty][t] * pj[tx][t];
  Ag[(size_t)(r0 + ty) * npad + c0 + tx] -= s;
}

// inverse of each NBxNB lower-triangular diagonal block of L -> diagonal blocks of W
__global__ void __launch_bounds__(NB) k_tri_diag_inv(const double* __restrict__ A, double* __restrict__ W, int npad) {
  __shared__ double l[NB][NB + 1];
  const int kb = blockIdx.x, g = blockIdx.y;
  const double* Ag = A + (size_t)g * npad * npad;
  double* Wg = W + (size_t)g * npad * npad;
  const int o = kb * NB;
  for (int e = threadIdx.x; e < NB * NB; e += NB) l[e / NB][e % NB] = Ag[(size_t)(o + e / NB) * npad + o + e % NB];
  __syncthreads();
  const int c = threadIdx.x;   // column of the inverse: solve L x = e_c
  double x[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
    for (int t = 0; t < r; ++t) s -= l[r][t] * x[t];
    x[r] = (r < c) ? 0.0 : s / l[r][r];
  }
#pragma unroll
  for (int r = 0; r < NB; ++r) Wg[(size_t)(o + r) * npad + o + c] = x[r];
}

// column block k of W = L^-1 by block forward substitution (one CTA per column block):
//   W_ik = -W_ii * sum_{j=k}^{i-1} L_ij W_jk ,  i = k+1 .. nblk-1
__global__ void __launch_bounds__(NB * NB) k_tri_inv_cols(const double* __restrict__ A, double* __restrict__ W, int npad) {
  __shared__ double sa[NB][NB + 1], sb[NB][NB + 1];
  const int k = blockIdx.x, g = blockIdx.y;
  const int nblk = npad / NB;
  const double* Ag = A + (size_t)g * npad * npad;
  double* Wg = W + (size_t)g * npad * npad;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = k + 1; i < nblk; ++i) {
    double t = 0.0;
    for (int j = k; j < i; ++j) {
      sa[ty][tx] = Ag[(size_t)(i * NB + ty) * npad + j * NB + tx];   // L_ij
      sb[ty][tx] = Wg[(size_t)(j * NB + ty) * npad + k * NB + tx];   // W_jk (written by this CTA earlier)
      __syncthreads();
#pragma unroll
      for (int q = 0; q < NB; ++q) t += sa[ty][q] * sb[q][tx];
      __syncthreads();
    }
    sa[ty][tx] = Wg[(size_t)(i * NB + ty) * npad + i * NB + tx];     // W_ii
    sb[ty][tx] = t;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int q = 0; q < NB; ++q) r += sa[ty][q] * sb[q][tx];
    Wg[(size_t)(i * NB + ty) * npad + k * NB + tx] = -r;
    __syncthreads();   // make W_ik visible to the whole CTA before the next block row reads it
  }
}

// alpha = W^T (W (y - m0))
__global__ void __launch_bounds__(256) k_alpha(ModelSpec ms, const double* __restrict__ Yn, double* __restrict__ tmp,
                                               double* __restrict__ alpha) {
  const int g = blockIdx.x;
  const int np = ms.npad;
  const double* Wg = ms.W + (size_t)g * np * np;
  double* t = tmp + (size_t)g * np;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int r = warp; r < np; r += nwarp) {   // one warp per row: coalesced reads of W[r][0..r]
    double s = 0.0;
    for (int c = lane; c <= r && c < ms.n; c += 32) s += Wg[(size_t)r * np + c] * (Yn[(size_t)c * ms.G + g] - ms.m0[g]);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) t[r] = (r < ms.n) ? s : 0.0;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < np; c += blockDim.x) {
    double s = 0.0;
    for (int r = c; r < ms.n; ++r) s += Wg[(size_t)r * np + c] * t[r];
    alpha[(size_t)g * np + c] = (c < ms.n) ? s : 0.0;
  }
}

int model_upload(sbo_ctx* ctx, int n, int d, int G, const double* X_norm, const double* Y_norm,
                 const double* X_mean, const double* X_std, const double* Y_mean, const double* Y_std,
                 const double* hyp) {
  SBO_REQUIRE(n >= 1 && n <= 16384, "n out of range");
  SBO_REQUIRE(d >= 1 && d <= SBO_MAX_D, "d out of range (1..8)");
  SBO_REQUIRE(G >= 1 && G <= SBO_MAX_G, "G out of range (1..8)");
  SBO_REQUIRE(X_norm && Y_norm && X_mean && X_std && Y_mean && Y_std && hyp, "null model pointer");
  ev_reset(ctx, 0);
  ev_begin(ctx, 0);
  ModelSpec& ms = ctx->ms;
  memset(&ms, 0, sizeof(ms));
  ms.n = n; ms.d = d; ms.G = G;
  ms.npad = (int)(cdiv(n, 128) * 128);   // multiple of the 128-row blocks of the DMMA solve kernel
  const int np = ms.npad;
  for (int k = 0; k < d; ++k) { ms.Xmean[k] = X_mean[k]; ms.Xstd[k] = X_std[k]; }
  for (int g = 0; g < G; ++g) {
    ms.Ymean[g] = Y_mean[g]; ms.Ystd[g] = Y_std[g];
    ms.m0[g] = (g == 0) ? 0.0 : (-2.0 * Y_mean[g]) / Y_std[g];                 // GP_Safe.py:331-332
    for (int k = 0; k < d; ++k) ms.inv_ell[g][k] = 1.0 / exp(2.0 * hyp[(size_t)k * G + g]);   // :338
    ms.sf2[g] = exp(2.0 * hyp[(size_t)d * G + g]);
    ms.sn2[g] = exp(2.0 * hyp[(size_t)(d + 1) * G + g]) + SBO_EPS_F32;         // :229
  }
  SBO_TRY(sbo_ensure(ctx, ctx->Xn, sizeof(double) * np * d));
  SBO_TRY(sbo_ensure(ctx, ctx->Yn, sizeof(double) * (size_t)n * G));
  SBO_TRY(sbo_ensure(ctx, ctx->alpha, sizeof(double) * (size_t)G * np * 2));
  SBO_TRY(sbo_ensure(ctx, ctx->W, sizeof(double) * (size_t)G * np * np));
  SBO_TRY(sbo_ensure(ctx, ctx->Kmat, sizeof(double) * (size_t)G * np * np));
  SBO_TRY(sbo_ensure(ctx, ctx->info, sizeof(int) * SBO_MAX_G));
  SBO_CUDA(cudaMemsetAsync(ctx->Xn.p, 0, sizeof(double) * np * d, ctx->stream));
  SBO_CUDA(cudaMemcpyAsync(ctx->Xn.p, X_norm, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, ctx->stream));
  SBO_CUDA(cudaMemcpyAsync(ctx->Yn.p, Y_norm, sizeof(double) * (size_t)n * G, cudaMemcpyHostToDevice, ctx->stream));
  SBO_CUDA(cudaMemsetAsync(ctx->W.p, 0, sizeof(double) * (size_t)G * np * np, ctx->stream));
  SBO_CUDA(cudaMemsetAsync(ctx->info.p, 0, sizeof(int) * SBO_MAX_G, ctx->stream));
  ms.Xn = (const double*)ctx->Xn.p;
  ms.alpha = (const double*)ctx->alpha.p;
  ms.W = (const double*)ctx->W.p;
  double* K = (double*)ctx->Kmat.p;
  double* W = (double*)ctx->W.p;

  {
    dim3 b(16, 16), g((unsigned)cdiv(np, 16), (unsigned)cdiv(np, 16), (unsigned)G);
    k_build_K<<<g, b, 0, ctx->stream>>>(ms, K);
    SBO_LAUNCH_CHECK();
  }
  const int nblk = np / NB;
  for (int kb = 0; kb < nblk; ++kb) {
    k_chol_diag<<<G, dim3(NB, NB), 0, ctx->stream>>>(K, np, kb, (int*)ctx->info.p);
    SBO_LAUNCH_CHECK();
    const int below = np - (kb + 1) * NB;
    if (below > 0) {
      k_chol_panel<<<dim3((unsigned)cdiv(below, 128), (unsigned)G), 128, 0, ctx->stream>>>(K, np, kb);
      SBO_LAUNCH_CHECK();
      const int nb = below / NB;
      k_chol_update<<<dim3(nb, nb, G), dim3(NB, NB), 0, ctx->stream>>>(K, np, kb);
      SBO_LAUNCH_CHECK();
    }
  }
  k_tri_diag_inv<<<dim3(nblk, G), NB, 0, ctx->stream>>>(K, W, np);
  SBO_LAUNCH_CHECK();
  k_tri_inv_cols<<<dim3(nblk, G), dim3(NB, NB), 0, ctx->stream>>>(K, W, np);
  SBO_LAUNCH_CHECK();
  k_alpha<<<G, 256, 0, ctx->stream>>>(ms, (const double*)ctx->Yn.p, (double*)ctx->alpha.p + (size_t)G * np,
                                       (double*)ctx->alpha.p);
  SBO_LAUNCH_CHECK();
  int info[SBO_MAX_G];
  SBO_CUDA(cudaMemcpyAsync(info, ctx->info.p, sizeof(int) * SBO_MAX_G, cudaMemcpyDeviceToHost, ctx->stream));
  ev_end(ctx);
  SBO_CUDA(cudaStreamSynchronize(ctx->stream));
  ev_collect(ctx);
  for (int g = 0; g < G; ++g)
    if (info[g] != 0)
      return sbo_fail(ctx, SBO_ERR_NUMERIC, "K of GP " + std::to_string(g) + " is not positive definite at pivot " +
                                                std::to_string(info[g] - 1));
  ctx->have_model = true;
  ctx->have_post = ctx->have_sets = ctx->have_sets2 = false;
  return SBO_OK;
}
