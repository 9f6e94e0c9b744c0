// pfh_eval_points_dev: P1 point evaluation of one field at a list of points (dolfin/pfbase.py:50-94, peval / sample).
#include <math_constants.h>

#include "pfh_internal.h"
#include "pfh_sample.cuh"

namespace {
__global__ void k_eval2(Mesh2 m, const double* __restrict__ u, const double* __restrict__ pts, long long npts, double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= npts) return;
  double v;
  out[i] = smp::eval2(m, u, pts[2 * i], pts[2 * i + 1], v) ? v : CUDART_INF;
}

// every rank writes the value for the cubes it owns, 0 for the others (summed across ranks), +inf outside the box
__global__ void k_eval3(Mesh3 m, FieldRef f, const double* __restrict__ pts, long long npts, double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= npts) return;
  bool inside, owned;
  double v = 0.0;
  smp::eval3(m, f, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], inside, owned, v);
  out[i] = !inside ? CUDART_INF : (owned ? v : 0.0);
}
}  // namespace

void k_eval_points(pfh_ctx* c, FieldRef f, const double* pts, long long npts, double* out) {
  if (npts <= 0) return;
  ProfScope ps(c, T_OTHER, 8.0 * (c->dim + 2) * npts);
  const int tpb = 128;
  const unsigned grid = (unsigned)((npts + tpb - 1) / tpb);
  if (c->dim == 3) k_eval3<<<grid, tpb, 0, c->stream>>>(c->m3, f, pts, npts, out);
  else k_eval2<<<grid, tpb, 0, c->stream>>>(c->m2, f.p, pts, npts, out);
  PFH_CUDA(cudaGetLastError());
}
