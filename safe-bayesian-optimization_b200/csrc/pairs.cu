// pairs.cu -- expander / GoOSE-target pair kernels.
//   Lipschitz mode (reference-exact):  ucb_idx(x) - L*||x - z + 1e-8||_2 >= 0, x in S, z in Z
//       models/SafeOpt.py:85-124 (Expander), models/GoOSE.py:69-114 (Target)
//   Fantasy mode, FP64 reference kernel (north_star):
//       c_i = k_i(z,x) - v_z.v_x ; rank-1 update of every constraint GP with y_i = ucb_i(x);
//       z newly safe iff every updated lcb_i(z) >= 0 ; g(x) = #newly safe z.
//   The TF32 tcgen05/TMEM version of the fantasy GEMM lives in fantasy_tc.cu.
#include "common.cuh"
#include <math.h>

#define PT 256   // threads per pair CTA = tile length of the staged side

struct PairConsts {
  int nc;                       // number of constraints G-1
  double L[SBO_MAX_G];          // L[c] for constraint c+1
  double beta;
};

// ---------------------------------------------------------------------------------------------
// payload gathers
// ---------------------------------------------------------------------------------------------
// coords[k][t] (SoA), ucb[c][t] for candidates t (local indices idx[t]); thr[c][t] = (ucb/L)^2 reach radius^2
__global__ void __launch_bounds__(256)
k_gather_points(GridSpec gs, int G, const long long* __restrict__ idx, long long n, const double* __restrict__ mean,
                const double* __restrict__ var, PairConsts pc, double* __restrict__ coords, double* __restrict__ ucb,
                double* __restrict__ thr) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long p = idx[t];
  double x[SBO_MAX_D];
  point_coords(gs, gs.first + p, x);
  for (int k = 0; k < gs.d; ++k) coords[(size_t)k * n + t] = x[k];
  if (ucb) {
    for (int c = 0; c < pc.nc; ++c) {
      const double u = ucb_of(mean[(size_t)(c + 1) * gs.count + p], var[(size_t)(c + 1) * gs.count + p], pc.beta);
      ucb[(size_t)c * n + t] = u;
      double r2;
      if (!(u >= 0.0)) r2 = -1.0;
      else if (pc.L[c] > 0.0) { const double r = u / pc.L[c]; r2 = r * r; }
      else r2 = INFINITY;
      thr[(size_t)c * n + t] = r2;
    }
  }
}

__device__ __forceinline__ bool reach_test(double s, double r2, double u, double L) {
  // exact reference predicate  u - L*sqrt(s) >= 0  (SafeOpt.py:85-88); the squared compare only
  // short-cuts pairs that are far (1e-12 relative) from the threshold.
  if (s <= r2 * (1.0 - 1e-12)) return true;
  if (s > r2 * (1.0 + 1e-12)) return false;
  return __dsub_rn(u, __dmul_rn(L, sqrt(s))) >= 0.0;
}

// ---------------------------------------------------------------------------------------------
// SafeOpt expander: one thread per candidate x, z tiles staged in shared memory.
// hits[c][t] = 1 if some z is reachable from x_t under constraint c+1.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(PT)
k_pairs_expander(PairConsts pc, long long nx, long long nz, const double* __restrict__ xc, const double* __restrict__ ucb,
                 const double* __restrict__ thr, const double* __restrict__ zc, unsigned char* __restrict__ hits,
                 unsigned long long* __restrict__ pair_counter, long long z_per_split) {
  __shared__ double zs[D][PT];
  const long long t = (long long)blockIdx.x * PT + threadIdx.x;
  const bool active = t < nx;
  double x[D], u[SBO_MAX_G - 1], r2[SBO_MAX_G - 1];
  const unsigned full = (1u << pc.nc) - 1u;
  unsigned found = 0;
  if (active) {
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = xc[(size_t)k * nx + t];
    for (int c = 0; c < pc.nc; ++c) {
      u[c] = ucb[(size_t)c * nx + t]; r2[c] = thr[(size_t)c * nx + t];
    }
  } else {
    found = full;
  }
  unsigned dead = 0;   // constraints that can never hit for this x
  if (active) for (int c = 0; c < pc.nc; ++c) if (r2[c] < 0.0) dead |= 1u << c;
  const long long z0 = (long long)blockIdx.y * z_per_split;
  const long long z1 = min(nz, z0 + z_per_split);
  unsigned long long tiles = 0;
  for (long long zb = z0; zb < z1; zb += PT) {
    if (__syncthreads_and((found | dead) == full)) break;
    const long long zi = zb + threadIdx.x;
#pragma unroll
    for (int k = 0; k < D; ++k) zs[k][threadIdx.x] = (zi < z1) ? zc[(size_t)k * nz + zi] : INFINITY;
    __syncthreads();
    ++tiles;
    if (active && (found | dead) != full) {
      const int tn = (int)min((long long)PT, z1 - zb);
      for (int j = 0; j < tn; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double df = __dadd_rn(__dsub_rn(x[k], zs[k][j]), SBO_PAIR_OFFSET);   // (x - z) + 1e-8, SafeOpt.py:87
          s = __dadd_rn(s, __dmul_rn(df, df));
        }
        for (int c = 0; c < pc.nc; ++c)
          if (!((found >> c) & 1u) && r2[c] >= 0.0 && reach_test(s, r2[c], u[c], pc.L[c])) found |= 1u << c;
      }
    }
  }
  if (active) {
    for (int c = 0; c < pc.nc; ++c)
      if ((found >> c) & 1u) hits[(size_t)c * nx + t] = 1;
  }
  if (threadIdx.x == 0 && pair_counter) atomicAdd(pair_counter, tiles * (unsigned long long)PT * PT * pc.nc);
}

// ---------------------------------------------------------------------------------------------
// GoOSE target: one thread per unsafe z, x tiles (coords, radius^2, ucb) staged in shared memory.
// hits[c][t] = 1 if z_t is reachable from some safe x under constraint c+1.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(PT)
k_pairs_target(PairConsts pc, long long nx, long long nz, const double* __restrict__ xc, const double* __restrict__ ucb,
               const double* __restrict__ thr, const double* __restrict__ zc, unsigned char* __restrict__ hits,
               unsigned long long* __restrict__ pair_counter, long long x_per_split) {
  __shared__ double xs[D][PT];
  __shared__ double us[SBO_MAX_G - 1][PT];
  __shared__ double rs[SBO_MAX_G - 1][PT];
  const long long t = (long long)blockIdx.x * PT + threadIdx.x;
  const bool active = t < nz;
  double z[D];
  const unsigned full = (1u << pc.nc) - 1u;
  unsigned found = active ? 0u : full;
  if (active) {
#pragma unroll
    for (int k = 0; k < D; ++k) z[k] = zc[(size_t)k * nz + t];
  }
  const long long x0 = (long long)blockIdx.y * x_per_split;
  const long long x1 = min(nx, x0 + x_per_split);
  unsigned long long tiles = 0;
  for (long long xb = x0; xb < x1; xb += PT) {
    if (__syncthreads_and(found == full)) break;
    const long long xi = xb + threadIdx.x;
#pragma unroll
    for (int k = 0; k < D; ++k) xs[k][threadIdx.x] = (xi < x1) ? xc[(size_t)k * nx + xi] : INFINITY;
    for (int c = 0; c < pc.nc; ++c) {
      us[c][threadIdx.x] = (xi < x1) ? ucb[(size_t)c * nx + xi] : -1.0;
      rs[c][threadIdx.x] = (xi < x1) ? thr[(size_t)c * nx + xi] : -1.0;
    }
    __syncthreads();
    ++tiles;
    if (active && found != full) {
      const int tn = (int)min((long long)PT, x1 - xb);
      for (int j = 0; j < tn; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          const double df = __dadd_rn(__dsub_rn(xs[k][j], z[k]), SBO_PAIR_OFFSET);   // (x - z) + 1e-8, GoOSE.py:71
          s = __dadd_rn(s, __dmul_rn(df, df));
        }
        for (int c = 0; c < pc.nc; ++c) {
          const double r2 = rs[c][j];
          if (!((found >> c) & 1u) && r2 >= 0.0 && reach_test(s, r2, us[c][j], pc.L[c])) found |= 1u << c;
        }
      }
    }
  }
  if (active) {
    for (int c = 0; c < pc.nc; ++c)
      if ((found >> c) & 1u) hits[(size_t)c * nz + t] = 1;
  }
  if (threadIdx.x == 0 && pair_counter) atomicAdd(pair_counter, tiles * (unsigned long long)PT * PT * pc.nc);
}

// hits[c][t] (per compacted element) -> bitmask words of the local shard, one mask per constraint
__global__ void __launch_bounds__(256)
k_hits_to_mask(int nc, long long n, const long long* __restrict__ idx, const unsigned char* __restrict__ hits,
               uint32_t* __restrict__ masks, long long nwords) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long p = idx[t];
  for (int c = 0; c < nc; ++c)
    if (hits[(size_t)c * n + t]) atomicOr(masks + (size_t)c * nwords + (p >> 5), 1u << (p & 31));
}
__global__ void __launch_bounds__(256)
k_union_count(int nc, const uint32_t* __restrict__ masks, long long nwords, unsigned long long* __restrict__ out) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned c = 0;
  if (w < nwords) {
    uint32_t m = 0;
    for (int i = 0; i < nc; ++i) m |= masks[(size_t)i * nwords + w];
    c = __popc(m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

template <int D>
static void launch_pairs(sbo_ctx* ctx, bool goose, const PairConsts& pc, long long nx, long long nz, const double* xc,
                         const double* ucb, const double* thr, const double* zc, unsigned char* hits,
                         unsigned long long* ctr) {
  const long long nthr = goose ? nz : nx, ntile = goose ? nx : nz;
  const long long bx = cdiv(nthr, PT);
  long long splits = 1;
  const long long tiles = cdiv(ntile, PT);
  while (bx * splits < 4 * 148 && splits * 8 <= tiles && splits < 65535) splits *= 2;   // enough CTAs, >= 8 tiles each
  const long long per = cdiv(cdiv(ntile, splits), PT) * PT;
  dim3 grid((unsigned)bx, (unsigned)cdiv(ntile, per));
  if (goose)
    k_pairs_target<D><<<grid, PT, 0, ctx->stream>>>(pc, nx, nz, xc, ucb, thr, zc, hits, ctr, per);
  else
    k_pairs_expander<D><<<grid, PT, 0, ctx->stream>>>(pc, nx, nz, xc, ucb, thr, zc, hits, ctr, per);
}

int pairs_lipschitz(sbo_ctx* ctx, bool goose, double beta, const double* L, sbo_pair_result* out) {
  SBO_REQUIRE(ctx->have_sets, "pair kernels need the sets (call sbo_sets)");
  SBO_REQUIRE(L != nullptr, "Lipschitz constants required");
  SBO_REQUIRE(out != nullptr, "null result");
  const ModelSpec& ms = ctx->ms;
  const GridSpec& gs = ctx->gs;
  const int nc = ms.G - 1;
  const long long count = gs.count, nw = mask_words(ctx);
  memset(out, 0, sizeof(*out));
  out->best_idx = -1;
  out->best_value = goose ? INFINITY : -INFINITY;
  for (int c = 0; c < SBO_MAX_G; ++c) { out->per_idx[c] = -1; out->per_value[c] = goose ? INFINITY : -INFINITY; }
  DevBuf& mbuf = goose ? ctx->m_tgt : ctx->m_exp;
  SBO_TRY(sbo_ensure(ctx, mbuf, sizeof(uint32_t) * (size_t)(nc > 0 ? nc : 1) * nw));
  SBO_CUDA(cudaMemsetAsync(mbuf.p, 0, sizeof(uint32_t) * (size_t)(nc > 0 ? nc : 1) * nw, ctx->stream));
  if (nc == 0) return SBO_OK;
  PairConsts pc{};
  pc.nc = nc; pc.beta = beta;
  for (int c = 0; c < nc; ++c) pc.L[c] = L[c + 1];

  ev_reset(ctx, 4); ev_reset(ctx, 6);
  ev_begin(ctx, 6);
  long long nx = 0, nz = 0;
  SBO_TRY(compact_mask(ctx, (const uint32_t*)ctx->m_safe.p, count, ctx->xs_idx, &nx));
  SBO_TRY(compact_mask(ctx, (const uint32_t*)ctx->m_unsafe.p, count, ctx->zs_idx, &nz));
  out->n_x = nx; out->n_z = nz;
  out->pairs_algorithmic = nx * nz * nc;
  if (nx == 0 || nz == 0) { ev_end(ctx); SBO_CUDA(cudaStreamSynchronize(ctx->stream)); ev_collect(ctx); return SBO_OK; }
  const int d = gs.d;
  SBO_TRY(sbo_ensure(ctx, ctx->xs_pay, sizeof(double) * (size_t)nx * (d + 2 * nc)));
  SBO_TRY(sbo_ensure(ctx, ctx->zs_pay, sizeof(double) * (size_t)nz * d));
  const long long nh = goose ? nz : nx;
  SBO_TRY(sbo_ensure(ctx, ctx->hits, (size_t)nc * nh));
  SBO_TRY(sbo_ensure(ctx, ctx->pairctr, 2 * sizeof(unsigned long long)));
  SBO_CUDA(cudaMemsetAsync(ctx->hits.p, 0, (size_t)nc * nh, ctx->stream));
  SBO_CUDA(cudaMemsetAsync(ctx->pairctr.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
  double* xc = (double*)ctx->xs_pay.p;
  double* ucb = xc + (size_t)d * nx;
  double* thr = ucb + (size_t)nc * nx;
  double* zc = (double*)ctx->zs_pay.p;
  k_gather_points<<<(unsigned)cdiv(nx, 256), 256, 0, ctx->stream>>>(gs, ms.G, (const long long*)ctx->xs_idx.p, nx,
                                                                    (const double*)ctx->mean.p, (const double*)ctx->var.p,
                                                                    pc, xc, ucb, thr);
  SBO_LAUNCH_CHECK();
  k_gather_points<<<(unsigned)cdiv(nz, 256), 256, 0, ctx->stream>>>(gs, ms.G, (const long long*)ctx->zs_idx.p, nz,
                                                                    (const double*)ctx->mean.p, (const double*)ctx->var.p,
                                                                    pc, zc, nullptr, nullptr);
  SBO_LAUNCH_CHECK();
  ev_end(ctx);
  unsigned long long* ctr = (unsigned long long*)ctx->pairctr.p;
  unsigned char* hits = (unsigned char*)ctx->hits.p;
  ev_begin(ctx, 4);
  switch (d) {
    case 1: launch_pairs<1>(ctx, goose, pc, nx, nz, xc, ucb, thr, zc, hits, ctr); break;
    case 2: launch_pairs<2>(ctx, goose, pc, nx, nz, xc, ucb, thr, zc, hits, ctr); break;
    case 3: launch_pairs<3>(ctx, goose, pc, nx, nz, xc, ucb, thr, zc, hits, ctr); break;
    case 4: launch_pairs<4>(ctx, goose, pc, nx, nz, xc, ucb, thr, zc, hits, ctr); break;
    case 5: launch_pairs<5>(ctx, goose, pc, nx, nz, xc, ucb, thr, zc, hits, ctr); break;
    case 6: launch_pairs<6>(ctx, goose, pc, nx, nz, xc, ucb, thr, zc, hits, ctr); break;
    case 7: launch_pairs<7>(ctx, goose, pc, nx, nz, xc, ucb, thr, zc, hits, ctr); break;
    default: launch_pairs<8>(ctx, goose, pc, nx, nz, xc, ucb, thr, zc, hits, ctr); break;
  }
  SBO_LAUNCH_CHECK();
  ev_end(ctx);
  ev_begin(ctx, 6);
  k_hits_to_mask<<<(unsigned)cdiv(nh, 256), 256, 0, ctx->stream>>>(nc, nh, (const long long*)(goose ? ctx->zs_idx.p : ctx->xs_idx.p),
                                                                  hits, (uint32_t*)mbuf.p, nw);
  SBO_LAUNCH_CHECK();
  k_union_count<<<(unsigned)cdiv(nw, 256), 256, 0, ctx->stream>>>(nc, (const uint32_t*)mbuf.p, nw, ctr + 1);
  SBO_LAUNCH_CHECK();
  ev_end(ctx);
  unsigned long long h[2];
  SBO_CUDA(cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  SBO_CUDA(cudaStreamSynchronize(ctx->stream));
  ev_collect(ctx);
  out->pairs_evaluated = (int64_t)h[0] < out->pairs_algorithmic ? (int64_t)h[0] : out->pairs_algorithmic;
  out->n_hit = (int64_t)h[1];
  // per-constraint arg-reduction, then first-best across constraints (SafeOpt.py:120-122 / GoOSE.py:110-112)
  const double keep5 = ctx->phase_ms[5];
  double acc5 = 0.0;
  for (int c = 0; c < nc; ++c) {
    int64_t idx; double val;
    SBO_TRY(argreduce_run(ctx, goose ? SBO_ARGMIN_LCB0 : SBO_ARGMAX_VAR0, (const uint32_t*)mbuf.p + (size_t)c * nw, nullptr, &idx, &val));
    acc5 += ctx->phase_ms[5];
    out->per_idx[c] = idx; out->per_value[c] = val;
    if (idx >= 0) {
      const bool better = out->best_idx < 0 || (goose ? (val < out->best_value) : (val > out->best_value));
      if (better) { out->best_idx = idx; out->best_value = val; }
    }
  }
  ctx->phase_ms[5] = keep5 + acc5;
  return SBO_OK;
}

// =============================================================================================
// Fantasy mode, FP64 SIMT reference kernel
// =============================================================================================
// gather V rows of the compacted points: Vout[c][t][0..npad) = vall[c][idx[t]][0..npad)  (zero rows for padding)
template <typename T>
__global__ void __launch_bounds__(256)
k_gather_rows(int nc, int npad, long long n, long long npadrows, long long vcount, const long long* __restrict__ idx,
              const T* __restrict__ vall, T* __restrict__ vout) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int c = blockIdx.y;
  if (row >= npadrows) return;
  const int lane = threadIdx.x & 31;
  T* o = vout + ((size_t)c * npadrows + row) * npad;
  if (row < n) {
    const T* s = vall + ((size_t)c * vcount + idx[row]) * npad;
    for (int k = lane; k < npad; k += 32) o[k] = s[k];
  } else {
    for (int k = lane; k < npad; k += 32) o[k] = (T)0;
  }
}

// per-point auxiliaries in NORMALISED units (double):
//  x side: xn[k][t], a[c][t] = beta*sigma/(sigma^2+sn2), b[c][t] = 1/(sigma^2+sn2)
//  z side: zn[k][t], m[c][t] = mean_raw/Ystd,            s[c][t] = var_raw/Ystd^2
__global__ void __launch_bounds__(256)
k_fantasy_aux(GridSpec gs, ModelSpec ms, FantasyConsts fc, const long long* __restrict__ idx, long long n,
              const double* __restrict__ mean, const double* __restrict__ var, int is_x,
              double* __restrict__ xn, double* __restrict__ a, double* __restrict__ b) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long p = idx[t];
  double x[SBO_MAX_D];
  point_coords(gs, gs.first + p, x);
  for (int k = 0; k < gs.d; ++k) xn[(size_t)k * n + t] = (x[k] - ms.Xmean[k]) / ms.Xstd[k];
  for (int c = 0; c < fc.nc; ++c) {
    const double ys = ms.Ystd[c + 1];
    const double vn = var[(size_t)(c + 1) * gs.count + p] / (ys * ys);
    if (is_x) {
      const double den = vn + fc.sn2[c];
      a[(size_t)c * n + t] = fc.beta * sqrt(vn) / den;
      b[(size_t)c * n + t] = 1.0 / den;
    } else {
      a[(size_t)c * n + t] = mean[(size_t)(c + 1) * gs.count + p] / ys;
      b[(size_t)c * n + t] = vn;
    }
  }
}

#define FB 64
#define FK 16
template <int D>
__global__ void __launch_bounds__(256)
k_fantasy_f64(FantasyConsts fc, long long nx, long long nz, long long nxp, long long nzp,
              const double* __restrict__ Vx, const double* __restrict__ Vz,
              const double* __restrict__ xn, const double* __restrict__ ax, const double* __restrict__ bx,
              const double* __restrict__ zn, const double* __restrict__ mz, const double* __restrict__ sz,
              int* __restrict__ counts) {
  __shared__ double As[FK][FB + 1];   // z rows
  __shared__ double Bs[FK][FB + 1];   // x rows
  __shared__ int cnt[16][FB];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long zb = (long long)blockIdx.x * FB, xb = (long long)blockIdx.y * FB;
  const int np = fc.npad;
  unsigned okmask = 0xffffu;   // bit i*4+j
  double zc[4][D], xc[4][D];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long zi = zb + ty * 4 + i, xi = xb + tx * 4 + i;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      zc[i][k] = (zi < nz) ? zn[(size_t)k * nz + zi] : 0.0;
      xc[i][k] = (xi < nx) ? xn[(size_t)k * nx + xi] : 0.0;
    }
  }
  for (int c = 0; c < fc.nc; ++c) {
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    const double* Az = Vz + ((size_t)c * nzp + zb) * np;
    const double* Bx = Vx + ((size_t)c * nxp + xb) * np;
    for (int k0 = 0; k0 < np; k0 += FK) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = tid + 256 * q;
        const int r = e >> 4, kk = e & 15;
        As[kk][r] = Az[(size_t)r * np + k0 + kk];
        Bs[kk][r] = Bx[(size_t)r * np + k0 + kk];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < FK; ++kk) {
        double a4[4], b4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a4[i] = As[kk][ty * 4 + i]; b4[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a4[i], b4[j], acc[i][j]);
      }
      __syncthreads();
    }
    // epilogue for constraint c
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long zi = zb + ty * 4 + i;
      const double m_z = (zi < nz) ? mz[(size_t)c * nz + zi] : 0.0;
      const double s_z = (zi < nz) ? sz[(size_t)c * nz + zi] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long xi = xb + tx * 4 + j;
        const double a_x = (xi < nx) ? ax[(size_t)c * nx + xi] : 0.0;
        const double b_x = (xi < nx) ? bx[(size_t)c * nx + xi] : 0.0;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) { const double df = zc[i][k] - xc[j][k]; s += df * df * fc.inv_ell[c][k]; }
        const double cc = fc.sf2[c] * exp(-0.5 * s) - acc[i][j];
        const double mu = m_z + cc * a_x;
        const double s2 = s_z - cc * cc * b_x;
        const bool ok = (mu - fc.beta * sqrt(fmax(s2, 0.0))) >= 0.0;
        if (!ok) okmask &= ~(1u << (i * 4 + j));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int cj = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long zi = zb + ty * 4 + i;
      if (zi < nz && ((okmask >> (i * 4 + j)) & 1u)) ++cj;
    }
    cnt[ty][tx * 4 + j] = cj;
  }
  __syncthreads();
  if (tid < FB) {
    int s = 0;
#pragma unroll
    for (int t = 0; t < 16; ++t) s += cnt[t][tid];
    const long long xi = xb + tid;
    if (xi < nx && s) atomicAdd(counts + xi, s);
  }
}

// counts per compacted candidate -> counts per local point + expander bitmask
__global__ void __launch_bounds__(256)
k_counts_scatter(long long nx, const long long* __restrict__ idx, const int* __restrict__ cnt_c, int* __restrict__ cnt_pt,
                 uint32_t* __restrict__ mask) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nx) return;
  const long long p = idx[t];
  const int c = cnt_c[t];
  if (cnt_pt) cnt_pt[p] = c;
  if (c > 0) atomicOr(mask + (p >> 5), 1u << (p & 31));
}

int fantasy_tc_run(sbo_ctx* ctx, const FantasyConsts& fc, int split, long long nx, long long nz, long long nxp,
                   long long nzp, const float* Vx, const float* Vz, const double* aux_x, const double* aux_z, int* counts_c);

int pairs_fantasy(sbo_ctx* ctx, int precision, double beta, sbo_pair_result* out, int32_t* counts_host) {
  SBO_REQUIRE(ctx->have_sets, "fantasy expander needs the sets (call sbo_sets)");
  SBO_REQUIRE(out != nullptr, "null result");
  SBO_REQUIRE(precision == SBO_PREC_FP64 || precision == SBO_PREC_TF32 || precision == SBO_PREC_TF32X3, "bad precision");
  const ModelSpec& ms = ctx->ms;
  const GridSpec& gs = ctx->gs;
  const int nc = ms.G - 1, d = gs.d, np = ms.npad;
  const long long count = gs.count, nw = mask_words(ctx);
  SBO_REQUIRE(nc >= 1, "fantasy expander needs at least one constraint GP");
  SBO_REQUIRE(ctx->keep_v == (precision == SBO_PREC_FP64 ? 1 : (precision == SBO_PREC_TF32 ? 2 : 3)),
              "fantasy expander: run sbo_posterior with keep_v = 1 (FP64), 2 (TF32) or 3 (TF32x3) first");
  const int rowlen = (precision == SBO_PREC_TF32X3) ? 2 * np : np;   // split rows are [hi | lo]
  memset(out, 0, sizeof(*out));
  out->best_idx = -1; out->best_value = -INFINITY;
  for (int c = 0; c < SBO_MAX_G; ++c) { out->per_idx[c] = -1; out->per_value[c] = -INFINITY; }
  SBO_TRY(sbo_ensure(ctx, ctx->m_exp, sizeof(uint32_t) * (size_t)nc * nw));
  SBO_CUDA(cudaMemsetAsync(ctx->m_exp.p, 0, sizeof(uint32_t) * (size_t)nc * nw, ctx->stream));
  SBO_TRY(sbo_ensure(ctx, ctx->counts, sizeof(int) * (size_t)count * 2));
  SBO_CUDA(cudaMemsetAsync(ctx->counts.p, 0, sizeof(int) * (size_t)count * 2, ctx->stream));
  FantasyConsts fc{};
  fc.nc = nc; fc.d = d; fc.npad = np; fc.beta = beta;
  for (int c = 0; c < nc; ++c) {
    fc.sf2[c] = ms.sf2[c + 1]; fc.sn2[c] = ms.sn2[c + 1];
    for (int k = 0; k < d; ++k) fc.inv_ell[c][k] = ms.inv_ell[c + 1][k];
  }
  ev_reset(ctx, 4); ev_reset(ctx, 6);
  ev_begin(ctx, 6);
  long long nx = 0, nz = 0;
  SBO_TRY(compact_mask(ctx, (const uint32_t*)ctx->m_safe.p, count, ctx->xs_idx, &nx));
  SBO_TRY(compact_mask(ctx, (const uint32_t*)ctx->m_unsafe.p, count, ctx->zs_idx, &nz));
  out->n_x = nx; out->n_z = nz;
  out->pairs_algorithmic = nx * nz * nc;
  out->pairs_evaluated = nx * nz * nc;
  if (nx == 0 || nz == 0) {
    ev_end(ctx); SBO_CUDA(cudaStreamSynchronize(ctx->stream)); ev_collect(ctx);
    if (counts_host) memset(counts_host, 0, sizeof(int32_t) * (size_t)count);
    return SBO_OK;
  }
  const long long nxp = cdiv(nx, 256) * 256, nzp = cdiv(nz, 256) * 256;
  const size_t esz = precision == SBO_PREC_FP64 ? sizeof(double) : sizeof(float);
  SBO_TRY(sbo_ensure(ctx, ctx->vx, esz * (size_t)nc * nxp * rowlen));
  SBO_TRY(sbo_ensure(ctx, ctx->vz, esz * (size_t)nc * nzp * rowlen));
  SBO_TRY(sbo_ensure(ctx, ctx->aux_x, sizeof(double) * (size_t)nx * (d + 2 * nc)));
  SBO_TRY(sbo_ensure(ctx, ctx->aux_z, sizeof(double) * (size_t)nz * (d + 2 * nc)));
  const long long* xi = (const long long*)ctx->xs_idx.p;
  const long long* zi = (const long long*)ctx->zs_idx.p;
  if (precision == SBO_PREC_FP64) {
    k_gather_rows<double><<<dim3((unsigned)cdiv(nxp, 8), nc), 256, 0, ctx->stream>>>(nc, np, nx, nxp, count, xi, (const double*)ctx->vall.p, (double*)ctx->vx.p);
    SBO_LAUNCH_CHECK();
    k_gather_rows<double><<<dim3((unsigned)cdiv(nzp, 8), nc), 256, 0, ctx->stream>>>(nc, np, nz, nzp, count, zi, (const double*)ctx->vall.p, (double*)ctx->vz.p);
    SBO_LAUNCH_CHECK();
  } else {
    k_gather_rows<float><<<dim3((unsigned)cdiv(nxp, 8), nc), 256, 0, ctx->stream>>>(nc, rowlen, nx, nxp, count, xi, (const float*)ctx->vall.p, (float*)ctx->vx.p);
    SBO_LAUNCH_CHECK();
    k_gather_rows<float><<<dim3((unsigned)cdiv(nzp, 8), nc), 256, 0, ctx->stream>>>(nc, rowlen, nz, nzp, count, zi, (const float*)ctx->vall.p, (float*)ctx->vz.p);
    SBO_LAUNCH_CHECK();
  }
  double* xn = (double*)ctx->aux_x.p; double* ax = xn + (size_t)d * nx; double* bx = ax + (size_t)nc * nx;
  double* zn = (double*)ctx->aux_z.p; double* mz = zn + (size_t)d * nz; double* sz = mz + (size_t)nc * nz;
  k_fantasy_aux<<<(unsigned)cdiv(nx, 256), 256, 0, ctx->stream>>>(gs, ms, fc, xi, nx, (const double*)ctx->mean.p, (const double*)ctx->var.p, 1, xn, ax, bx);
  SBO_LAUNCH_CHECK();
  k_fantasy_aux<<<(unsigned)cdiv(nz, 256), 256, 0, ctx->stream>>>(gs, ms, fc, zi, nz, (const double*)ctx->mean.p, (const double*)ctx->var.p, 0, zn, mz, sz);
  SBO_LAUNCH_CHECK();
  ev_end(ctx);
  int* cnt_c = (int*)ctx->counts.p;            // per compacted candidate
  int* cnt_pt = cnt_c + count;                 // per local point
  if (precision == SBO_PREC_FP64) {
    ev_begin(ctx, 4);
    dim3 grid((unsigned)cdiv(nz, FB), (unsigned)cdiv(nx, FB));
    SBO_REQUIRE(grid.y <= 65535, "too many candidates for the FP64 fantasy kernel");
#define FL(DD) k_fantasy_f64<DD><<<grid, 256, 0, ctx->stream>>>(fc, nx, nz, nxp, nzp, (const double*)ctx->vx.p, (const double*)ctx->vz.p, xn, ax, bx, zn, mz, sz, cnt_c)
    switch (d) { case 1: FL(1); break; case 2: FL(2); break; case 3: FL(3); break; case 4: FL(4); break;
                 case 5: FL(5); break; case 6: FL(6); break; case 7: FL(7); break; default: FL(8); break; }
#undef FL
    SBO_LAUNCH_CHECK();
  } else {   // record prep is logged as phase 6, the GEMM kernel as phase 4 (begun inside)
    SBO_TRY(fantasy_tc_run(ctx, fc, precision == SBO_PREC_TF32X3 ? 1 : 0, nx, nz, nxp, nzp, (const float*)ctx->vx.p, (const float*)ctx->vz.p,
                           (const double*)ctx->aux_x.p, (const double*)ctx->aux_z.p, cnt_c));
  }
  ev_end(ctx);
  ev_begin(ctx, 6);
  k_counts_scatter<<<(unsigned)cdiv(nx, 256), 256, 0, ctx->stream>>>(nx, xi, cnt_c, cnt_pt, (uint32_t*)ctx->m_exp.p);
  SBO_LAUNCH_CHECK();
  SBO_TRY(sbo_ensure(ctx, ctx->pairctr, 2 * sizeof(unsigned long long)));
  SBO_CUDA(cudaMemsetAsync(ctx->pairctr.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
  k_union_count<<<(unsigned)cdiv(nw, 256), 256, 0, ctx->stream>>>(1, (const uint32_t*)ctx->m_exp.p, nw, (unsigned long long*)ctx->pairctr.p + 1);
  SBO_LAUNCH_CHECK();
  ev_end(ctx);
  unsigned long long h[2];
  SBO_CUDA(cudaMemcpyAsync(h, ctx->pairctr.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  if (counts_host) SBO_CUDA(cudaMemcpyAsync(counts_host, cnt_pt, sizeof(int32_t) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  SBO_CUDA(cudaStreamSynchronize(ctx->stream));
  ev_collect(ctx);
  out->n_hit = (int64_t)h[1];
  int64_t idx; double val;
  SBO_TRY(argreduce_run(ctx, SBO_ARGMAX_VAR0, (const uint32_t*)ctx->m_exp.p, nullptr, &idx, &val));
  out->per_idx[0] = idx; out->per_value[0] = val;
  out->best_idx = idx; out->best_value = val;
  return SBO_OK;
}
