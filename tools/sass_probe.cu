// One field product per kernel, so that `cuobjdump -sass` shows exactly the instruction mix of the primitives every
// hot kernel inlines (tools/sass_evidence.sh writes the listings and the opcode counts to profiles/).
#include "../gemini_b200/csrc/fp.cuh"
using namespace gm;

__global__ void probe_fq_mul(const Fq* a, const Fq* b, Fq* r) { r[threadIdx.x] = a[threadIdx.x] * b[threadIdx.x]; }
__global__ void probe_fr_mul(const Fr* a, const Fr* b, Fr* r) { r[threadIdx.x] = a[threadIdx.x] * b[threadIdx.x]; }
__global__ void probe_fr_lazy_mul_add(const Fr* a, const Fr* b, Fr* r) {
  FrAcc acc = FrAcc::zero();
  acc.mul_add(a[threadIdx.x], b[threadIdx.x]);
  acc.mul_add(a[threadIdx.x + 32], b[threadIdx.x + 32]);
  r[threadIdx.x] = acc.reduce();
}
