// Self-test kernels: run the device field / curve primitives element-wise so that the GPU parity
// tests (tests/test_gpu_arith.py) can compare them with the big-integer oracle.
#include "common.cuh"
#include "g1.cuh"

namespace gm {

template <class F>
__global__ void k_selftest_field(int op, const F* a, const F* b, F* r, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = a[i], y = b[i], z;
  switch (op) {
    case 0: z = x * y; break;
    case 1: z = x + y; break;
    case 2: z = x - y; break;
    case 3: z = fp_inv(x); break;
    case 4: z = x.from_mont(); break;
    case 5: z = x.to_mont(); break;
    case 7: z = fp_inv_divsteps(x); break;
    default: z = x.sqr(); break;
  }
  r[i] = z;
}

__global__ void k_selftest_curve(int op, const XYZZ* acc, const uint32_t* other, Jacobian* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ a = acc[i];
  if (op == 0 || op == 1) {
    Affine p = reinterpret_cast<const Affine*>(other)[i];
    if (op == 1) p.y = p.y.neg();
    xyzz_madd(a, p);
  } else if (op == 2) {
    XYZZ b = reinterpret_cast<const XYZZ*>(other)[i];
    xyzz_add(a, b);
  } else {
    xyzz_dbl(a);
  }
  out[i] = xyzz_to_jacobian_normalized(a);
}

}  // namespace gm

using namespace gm;

extern "C" int gm_selftest_field(gm_ctx* ctx, int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* r, size_t n) {
  GM_ARG(ctx && a && b && r, "NULL argument");
  GM_ARG(field == 0 || field == 1, "field must be 0 (Fq) or 1 (Fr)");
  GM_ARG(op >= 0 && op <= 7, "unknown op");
  GM_TRY(set_device(ctx));
  const size_t bytes = n * (field == 0 ? 48 : 32);
  void *da, *db, *dr;
  GM_CUDA(cudaMalloc(&da, bytes + 16));
  GM_CUDA(cudaMalloc(&db, bytes + 16));
  GM_CUDA(cudaMalloc(&dr, bytes + 16));
  GM_CUDA(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
  GM_CUDA(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (n) {
    if (field == 0) k_selftest_field<Fq><<<grid, 128, 0, ctx->stream>>>(op, (const Fq*)da, (const Fq*)db, (Fq*)dr, n);
    else k_selftest_field<Fr><<<grid, 128, 0, ctx->stream>>>(op, (const Fr*)da, (const Fr*)db, (Fr*)dr, n);
    ctx->launches++;
  }
  GM_CUDA(cudaGetLastError());
  GM_CUDA(cudaMemcpyAsync(r, dr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(da); cudaFree(db); cudaFree(dr);
  return GM_OK;
}

extern "C" int gm_selftest_curve(gm_ctx* ctx, int op, const uint32_t* acc_xyzz, const uint32_t* other, uint32_t* out_jac, size_t n) {
  GM_ARG(ctx && acc_xyzz && out_jac && (other || op == 3), "NULL argument");
  GM_ARG(op >= 0 && op <= 3, "unknown op");
  GM_TRY(set_device(ctx));
  const size_t other_bytes = (op <= 1 ? 96 : 192) * n;
  void *da, *db, *dr;
  GM_CUDA(cudaMalloc(&da, n * 192 + 16));
  GM_CUDA(cudaMalloc(&db, other_bytes + 16));
  GM_CUDA(cudaMalloc(&dr, n * 144 + 16));
  GM_CUDA(cudaMemcpyAsync(da, acc_xyzz, n * 192, cudaMemcpyHostToDevice, ctx->stream));
  if (op != 3) GM_CUDA(cudaMemcpyAsync(db, other, other_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (n) {
    k_selftest_curve<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(op, (const XYZZ*)da, (const uint32_t*)db, (Jacobian*)dr, n);
    ctx->launches++;
  }
  GM_CUDA(cudaGetLastError());
  GM_CUDA(cudaMemcpyAsync(out_jac, dr, n * 144, cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(da); cudaFree(db); cudaFree(dr);
  return GM_OK;
}
