// Krylov vector kernels (fp64, HBM-streaming): fused multi-dot and multi-axpy for classical
// Gram-Schmidt GMRES, norms, axpy-type updates and the deterministic two-stage reductions.
// These replace PETSc's VecMDot / VecMAXPY / VecNorm / VecAXPY (+ MPI_Allreduce) on the
// reference's hot path (hot loop E).
#include "pfh_internal.h"

namespace {
constexpr int TPB = 256;
constexpr int MD = 8;  // basis vectors handled per multi-dot / multi-axpy pass (16 per pass measured slower: 2.6 vs 1.7 ms, the loads serialise at 64 registers)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

int grid_stream(long long n) {  // enough CTAs to fill 148 SMs x 8, grid-stride beyond that
  long long g = (n + TPB * 2 - 1) / (TPB * 2);
  if (g < 1) g = 1;
  return (int)(g < PFH_RED_MAXBLOCKS ? g : PFH_RED_MAXBLOCKS);
}

struct Coef {
  double v[MD];
};

// Basis vectors are stored in fp64 or, for the compressed-basis GMRES (pfh_solver.cpp), in fp32; arithmetic is fp64.
// Every thread moves 16-byte chunks of each basis vector: 2 doubles or 4 floats (same bytes in flight either way).
template <class TV>
struct Chunk;
template <>
struct Chunk<double> {
  static constexpr int E = 2;
  static __device__ __forceinline__ void load(const double* p, long long i, double (&v)[2]) {
    const double2 t = reinterpret_cast<const double2*>(p)[i];
    v[0] = t.x; v[1] = t.y;
  }
};
template <>
struct Chunk<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void load(const float* p, long long i, double (&v)[4]) {
    const float4 t = reinterpret_cast<const float4*>(p)[i];
    v[0] = (double)t.x; v[1] = (double)t.y; v[2] = (double)t.z; v[3] = (double)t.w;
  }
};
template <int E>
__device__ __forceinline__ void load_w(const double* p, long long i, double (&v)[E]) {
#pragma unroll
  for (int q = 0; q < E / 2; ++q) {
    const double2 t = reinterpret_cast<const double2*>(p)[i * (E / 2) + q];
    v[2 * q] = t.x; v[2 * q + 1] = t.y;
  }
}

// partial[block][k0 + k] = sum_i V_{k0+k}[i] w[i] (k < nb); if WW also partial[block][ww_slot] = sum w^2
template <bool WW, class TV>
__global__ void __launch_bounds__(TPB) k_multidot(const TV* __restrict__ V, long long stride, int nb, int k0, int ww_slot,
                                                  const double* __restrict__ w, long long n, double* __restrict__ partial) {
  constexpr int E = Chunk<TV>::E;
  double acc[MD + 1];
#pragma unroll
  for (int k = 0; k <= MD; ++k) acc[k] = 0.0;
  const long long nc = n / E;
  const long long step = (long long)gridDim.x * TPB;
  long long i = blockIdx.x * (long long)TPB + threadIdx.x;
  // two chunks per trip: twice the loads in flight per thread (with one or two basis vectors the kernel is latency-bound)
  for (; i + step < nc; i += 2 * step) {
    double wa[E], wb[E];
    load_w<E>(w, i, wa);
    load_w<E>(w, i + step, wb);
#pragma unroll
    for (int k = 0; k < MD; ++k)
      if (k < nb) {
        double va[E], vb[E];
        Chunk<TV>::load(V + (long long)k * stride, i, va);
        Chunk<TV>::load(V + (long long)k * stride, i + step, vb);
#pragma unroll
        for (int e = 0; e < E; ++e) acc[k] += va[e] * wa[e];
#pragma unroll
        for (int e = 0; e < E; ++e) acc[k] += vb[e] * wb[e];
      }
    if (WW) {
#pragma unroll
      for (int e = 0; e < E; ++e) acc[MD] += wa[e] * wa[e];
#pragma unroll
      for (int e = 0; e < E; ++e) acc[MD] += wb[e] * wb[e];
    }
  }
  for (; i < nc; i += step) {
    double wi[E];
    load_w<E>(w, i, wi);
#pragma unroll
    for (int k = 0; k < MD; ++k)
      if (k < nb) {
        double vi[E];
        Chunk<TV>::load(V + (long long)k * stride, i, vi);
#pragma unroll
        for (int e = 0; e < E; ++e) acc[k] += vi[e] * wi[e];
      }
    if (WW) {
#pragma unroll
      for (int e = 0; e < E; ++e) acc[MD] += wi[e] * wi[e];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long t = nc * E; t < n; ++t) {
      const double wi = w[t];
#pragma unroll
      for (int k = 0; k < MD; ++k)
        if (k < nb) acc[k] += (double)V[(long long)k * stride + t] * wi;
      if (WW) acc[MD] += wi * wi;
    }
  __shared__ double sh[MD + 1][TPB / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k <= MD; ++k) {
    const double s = warp_sum(acc[k]);
    if (lane == 0) sh[k][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x <= MD) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < TPB / 32; ++q) s += sh[threadIdx.x][q];
    if (threadIdx.x < nb) partial[blockIdx.x * PFH_RED_MAXVALS + k0 + threadIdx.x] = s;
    if (WW && threadIdx.x == MD) partial[blockIdx.x * PFH_RED_MAXVALS + ww_slot] = s;
  }
}

// out = wscale * w - sum_k h[k] V_k   (HAS_W false: out = - sum_k h[k] V_k)
// out32 != nullptr: the result is rounded to fp32 and stored there (the next compressed basis vector); `out` then
// receives that rounded value, so the fp64 working copy and the stored basis vector are the same vector.
// NORM: partial[block][0] = sum over this block's elements of (stored result)^2, for the norm of the new basis vector
template <bool HAS_W, class TV, bool NORM>
__global__ void __launch_bounds__(TPB) k_multiaxpy(const TV* __restrict__ V, long long stride, int nb, Coef h,
                                                   const double* w, double wscale, double* out, float* out32, long long n,
                                                   double* __restrict__ partial) {
  constexpr int E = Chunk<TV>::E;
  double nacc = 0.0;
  const long long nc = n / E;
  for (long long i = blockIdx.x * (long long)TPB + threadIdx.x; i < nc; i += (long long)gridDim.x * TPB) {
    double r[E];
    if (HAS_W) {
      load_w<E>(w, i, r);
#pragma unroll
      for (int e = 0; e < E; ++e) r[e] *= wscale;
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) r[e] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < MD; ++k)
      if (k < nb) {
        double vi[E];
        Chunk<TV>::load(V + (long long)k * stride, i, vi);
#pragma unroll
        for (int e = 0; e < E; ++e) r[e] -= h.v[k] * vi[e];
      }
    if (out32) {
#pragma unroll
      for (int q = 0; q < E / 2; ++q) {
        const float2 f = make_float2((float)r[2 * q], (float)r[2 * q + 1]);
        reinterpret_cast<float2*>(out32)[i * (E / 2) + q] = f;
        r[2 * q] = (double)f.x;
        r[2 * q + 1] = (double)f.y;
      }
    }
#pragma unroll
    for (int q = 0; q < E / 2; ++q) {
      if (NORM) nacc += r[2 * q] * r[2 * q] + r[2 * q + 1] * r[2 * q + 1];
      reinterpret_cast<double2*>(out)[i * (E / 2) + q] = make_double2(r[2 * q], r[2 * q + 1]);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long t = nc * E; t < n; ++t) {
      double r = HAS_W ? w[t] * wscale : 0.0;
#pragma unroll
      for (int k = 0; k < MD; ++k)
        if (k < nb) r -= h.v[k] * (double)V[(long long)k * stride + t];
      if (out32) {
        out32[t] = (float)r;
        r = (double)out32[t];
      }
      if (NORM) nacc += r * r;
      out[t] = r;
    }
  if (NORM) {
    __shared__ double sh[TPB / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    nacc = warp_sum(nacc);
    if (lane == 0) sh[wid] = nacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < TPB / 32; ++q) t += sh[q];
      partial[blockIdx.x * PFH_RED_MAXVALS] = t;
    }
  }
}

// out = a x + b y  (y may be null when b == 0)
__global__ void __launch_bounds__(TPB) k_axpby(double a, const double* __restrict__ x, double b, const double* y,
                                               double* out, long long n) {
  for (long long i = blockIdx.x * (long long)TPB + threadIdx.x; i < n; i += (long long)gridDim.x * TPB)
    out[i] = a * x[i] + (b != 0.0 ? b * y[i] : 0.0);
}

__global__ void __launch_bounds__(TPB) k_norm2(const double* __restrict__ x, long long n, double* __restrict__ partial) {
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)TPB + threadIdx.x; i < n; i += (long long)gridDim.x * TPB) acc += x[i] * x[i];
  __shared__ double sh[TPB / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  acc = warp_sum(acc);
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < TPB / 32; ++q) s += sh[q];
    partial[blockIdx.x * PFH_RED_MAXVALS] = s;
  }
}

// second stage: red[v] = sum_b partial[b][v], fixed order => deterministic.  One warp per value (8 at a time).  With
// `host` given (single rank) the results also go straight into mapped pinned host memory, followed by a sequence number
// the host polls: no device-to-host copy, no cudaStreamSynchronize on the Krylov iteration's critical path.
__global__ void __launch_bounds__(TPB) k_reduce_final(const double* __restrict__ partial, int nblocks, int nvals,
                                                      double* __restrict__ red, double* host, unsigned long long* host_seq,
                                                      unsigned long long seq) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int v = wid; v < nvals; v += TPB / 32) {
    double acc = 0.0;
    for (int b = lane; b < nblocks; b += 32) acc += partial[b * PFH_RED_MAXVALS + v];
    acc = warp_sum(acc);
    if (lane == 0) {
      red[v] = acc;
      if (host != nullptr) host[v] = acc;
    }
  }
  if (host != nullptr) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      *reinterpret_cast<volatile unsigned long long*>(host_seq) = seq;
      __threadfence_system();
    }
  }
}

// after an allreduce: the reduced values from device memory into the mapped host buffer + sequence number
__global__ void k_publish(const double* __restrict__ red, int nvals, double* host, unsigned long long* host_seq, unsigned long long seq) {
  if (threadIdx.x < nvals) host[threadIdx.x] = red[threadIdx.x];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned long long*>(host_seq) = seq;
    __threadfence_system();
  }
}
}  // namespace

const double* pfh_reduce_finish(pfh_ctx* c, int nblocks, int nvals) {
  PFH_CHECK(nvals <= PFH_RED_MAXVALS && nblocks <= PFH_RED_MAXBLOCKS, "reduction scratch too small");
  const bool multi = c->cfg.nranks > 1;
  const unsigned long long seq = ++c->red_seq;
  {
    ProfScope ps(c, T_REDUCE);
    k_reduce_final<<<1, TPB, 0, c->stream>>>(c->d_partial, nblocks, nvals, c->d_red, multi ? nullptr : c->h_red, c->h_seq, seq);
    PFH_CUDA(cudaGetLastError());
  }
  if (multi) {
    comm_allreduce_sum(c, c->d_red, nvals);
    k_publish<<<1, 64, 0, c->stream>>>(c->d_red, nvals, c->h_red, c->h_seq, seq);
    PFH_CUDA(cudaGetLastError());
  }
  // wait for the sequence number in host memory; the stream is queried now and then so that a failed launch surfaces
  const volatile unsigned long long* hs = c->h_seq;
  for (unsigned long long spins = 1;; ++spins) {
    if (*hs == seq) break;
    if ((spins & 0x3fff) == 0) {
      const cudaError_t e = cudaStreamQuery(c->stream);
      if (e == cudaSuccess) {
        if (*hs != seq) PFH_CUDA(cudaStreamSynchronize(c->stream));
        PFH_CHECK(*hs == seq, "internal: reduction result did not arrive in host memory");
        break;
      }
      if (e != cudaErrorNotReady) PFH_CUDA(e);
    }
  }
  return c->h_red;
}

void v_copy(pfh_ctx* c, double* dst, const double* src, long long n) {
  ProfScope ps(c, T_COPY, 16.0 * n);
  PFH_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
}
void v_zero(pfh_ctx* c, double* dst, long long n) {
  ProfScope ps(c, T_COPY, 8.0 * n);
  PFH_CUDA(cudaMemsetAsync(dst, 0, n * sizeof(double), c->stream));
}
void v_axpby(pfh_ctx* c, double a, const double* x, double b, const double* y, double* out, long long n) {
  ProfScope ps(c, T_VEC, 8.0 * (b != 0.0 ? 3 : 2) * n);
  k_axpby<<<grid_stream(n), TPB, 0, c->stream>>>(a, x, b, y, out, n);
  PFH_CUDA(cudaGetLastError());
}
void v_axpy(pfh_ctx* c, double a, const double* x, double* y, long long n) { v_axpby(c, a, x, 1.0, y, y, n); }
void v_scale(pfh_ctx* c, double a, const double* x, double* out, long long n) { v_axpby(c, a, x, 0.0, nullptr, out, n); }

template <class TV>
static void multidot_t(pfh_ctx* c, const TV* V, long long stride, int nv, const double* w, long long n, double* host_out) {
  PFH_CHECK(nv + 1 <= PFH_RED_MAXVALS, "too many basis vectors for the reduction scratch");
  PFH_CHECK((stride & 3) == 0 && (reinterpret_cast<uintptr_t>(V) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
            "multidot operands must be 16-byte aligned");
  const int nb = grid_stream(n);
  int k0 = 0;
  do {
    const int cnt = nv - k0 < MD ? nv - k0 : MD;
    const bool last = k0 + cnt >= nv;
    ProfScope ps(c, T_MULTIDOT, ((double)sizeof(TV) * cnt + 8.0) * n);
    if (last)
      k_multidot<true, TV><<<nb, TPB, 0, c->stream>>>(V + (long long)k0 * stride, stride, cnt, k0, nv, w, n, c->d_partial);
    else
      k_multidot<false, TV><<<nb, TPB, 0, c->stream>>>(V + (long long)k0 * stride, stride, cnt, k0, nv, w, n, c->d_partial);
    PFH_CUDA(cudaGetLastError());
    k0 += cnt;
  } while (k0 < nv);
  const double* h = pfh_reduce_finish(c, nb, nv + 1);
  for (int k = 0; k <= nv; ++k) host_out[k] = h[k];
}

// out = wscale * w - sum_k coef_k V_k, MD basis vectors per pass (w == nullptr: no w term); the last pass also
// writes the fp32-rounded result to out32 when given
template <class TV>
static void multiaxpy_raw(pfh_ctx* c, const TV* V, long long stride, int nv, const double* coef, const double* w,
                          double wscale, double* out, float* out32, long long n, double* norm2 = nullptr) {
  PFH_CHECK((stride & 3) == 0 && (reinterpret_cast<uintptr_t>(V) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(out32) & 7) == 0,
            "multiaxpy operands must be 16-byte aligned");
  int k0 = 0;
  bool first = true;
  do {
    const int cnt = nv - k0 < MD ? nv - k0 : MD;
    Coef cf;
    for (int k = 0; k < MD; ++k) cf.v[k] = k < cnt ? coef[k0 + k] : 0.0;
    const double* win = first ? w : out;
    const bool last = k0 + cnt >= nv;
    float* o32 = last ? out32 : nullptr;
    ProfScope ps(c, T_MULTIAXPY, ((double)sizeof(TV) * cnt + 8.0 * (win ? 2 : 1) + (o32 ? 4.0 : 0.0)) * n);
    const int nb = grid_stream(n);
    const TV* Vk = V + (long long)k0 * stride;
    if (last && norm2) {
      PFH_CHECK(win != nullptr, "multiaxpy: norm needs a w term");
      k_multiaxpy<true, TV, true><<<nb, TPB, 0, c->stream>>>(Vk, stride, cnt, cf, win, first ? wscale : 1.0, out, o32, n, c->d_partial);
    } else if (win) {
      k_multiaxpy<true, TV, false><<<nb, TPB, 0, c->stream>>>(Vk, stride, cnt, cf, win, first ? wscale : 1.0, out, o32, n, nullptr);
    } else {
      k_multiaxpy<false, TV, false><<<nb, TPB, 0, c->stream>>>(Vk, stride, cnt, cf, nullptr, 0.0, out, o32, n, nullptr);
    }
    PFH_CUDA(cudaGetLastError());
    first = false;
    k0 += cnt;
  } while (k0 < nv);
  if (out32 != nullptr) s_after_vector_written(c, out);
  if (norm2) *norm2 = pfh_reduce_finish(c, grid_stream(n), 1)[0];
}

void v_multidot(pfh_ctx* c, const double* V, long long stride, int nv, const double* w, long long n, double* host_out) {
  multidot_t(c, V, stride, nv, w, n, host_out);
}
void v_multidot32(pfh_ctx* c, const float* V, long long stride, int nv, const double* w, long long n, double* host_out) {
  multidot_t(c, V, stride, nv, w, n, host_out);
}

void v_multiaxpy(pfh_ctx* c, const double* V, long long stride, int nv, const double* h, const double* w, double scale,
                 double* out, long long n) {
  std::vector<double> cf(nv);
  for (int k = 0; k < nv; ++k) cf[k] = scale * h[k];
  multiaxpy_raw(c, V, stride, nv, cf.data(), w, scale, out, nullptr, n);
}
void v_multiaxpy32(pfh_ctx* c, const float* V, long long stride, int nv, const double* h, const double* w, double scale,
                   double* out, float* out32, long long n, double* norm2) {
  std::vector<double> cf(nv);
  for (int k = 0; k < nv; ++k) cf[k] = scale * h[k];
  multiaxpy_raw(c, V, stride, nv, cf.data(), w, scale, out, out32, n, norm2);
}

void v_combine(pfh_ctx* c, const double* V, long long stride, int nv, const double* y, double* out, long long n) {
  std::vector<double> cf(nv);
  for (int k = 0; k < nv; ++k) cf[k] = -y[k];
  multiaxpy_raw(c, V, stride, nv, cf.data(), nullptr, 0.0, out, nullptr, n);
}
void v_combine32(pfh_ctx* c, const float* V, long long stride, int nv, const double* y, double* out, long long n) {
  std::vector<double> cf(nv);
  for (int k = 0; k < nv; ++k) cf[k] = -y[k];
  multiaxpy_raw(c, V, stride, nv, cf.data(), nullptr, 0.0, out, nullptr, n);
}

double v_norm2(pfh_ctx* c, const double* x, long long n) {
  const int nb = grid_stream(n);
  {
    ProfScope ps(c, T_VEC, 8.0 * n);
    k_norm2<<<nb, TPB, 0, c->stream>>>(x, n, c->d_partial);
    PFH_CUDA(cudaGetLastError());
  }
  return pfh_reduce_finish(c, nb, 1)[0];
}
