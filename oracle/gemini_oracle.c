/* CPU restatement (plain C) of the reference's algorithms for the Gemini prover hot path.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ as a fast checker at sizes the pure-Python oracle
 * (oracle/pyref.py) cannot reach, and by bench.py's cpu_baseline / --impl reference legs as the
 * timed CPU baseline.  Never imported, linked or executed by the product (gemini_b200/).
 *
 * The reference (arkworks-rs/gemini @ 844a85e5, Rust) cannot be compiled here: no cargo/rustc in
 * the image and its arithmetic crates (ark-ec / ark-ff 0.4.2, Cargo.lock:44-46,62-64) are not
 * vendored.  This file restates, citing /root/reference paths:
 *   - go_msm_g1          VariableBaseMSM::msm_unchecked = into_bigint + signed-digit windowed
 *                        Pippenger: src/kzg/msm/variable_base.rs:16-19 (window size), :21-61 (digits),
 *                        :95-177 (buckets, running sum, window combination); one thread per window,
 *                        mirroring ark-ec's rayon task per window
 *   - go_fr_fold         misc::fold_polynomial, src/misc.rs:52-56 (single-threaded like the reference)
 *   - go_sumcheck_time   TimeProver::{fold,next_message,final_foldings},
 *                        src/subprotocols/sumcheck/time_prover.rs:75-137 (single-threaded)
 * Parity status: checked against oracle/pyref.py (naive double-and-add, reference KATs) by
 * tests/test_oracle_c.py.  For the MSM value the reference holds no golden vector: "parity unpinned"
 * at the ark-ec boundary; the oracle of record is the naive sum in pyref.py.
 *
 * Field elements: little-endian u64 limbs, Montgomery form (arkworks' in-memory form).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ---------------------------------------------------------------- generic Montgomery (N limbs) */
#define DEFINE_FIELD(PFX, N, MOD, INV, ONE)                                                     \
  static const uint64_t PFX##_mod[N] = MOD;                                                     \
  static const uint64_t PFX##_one[N] = ONE;                                                     \
  static inline int PFX##_is_zero(const uint64_t* a) {                                          \
    uint64_t x = 0;                                                                             \
    for (int i = 0; i < N; i++) x |= a[i];                                                      \
    return x == 0;                                                                              \
  }                                                                                             \
  static inline int PFX##_eq(const uint64_t* a, const uint64_t* b) {                            \
    uint64_t x = 0;                                                                             \
    for (int i = 0; i < N; i++) x |= a[i] ^ b[i];                                               \
    return x == 0;                                                                              \
  }                                                                                             \
  static inline void PFX##_csub(uint64_t* r, const uint64_t* t, uint64_t top) {                 \
    uint64_t s[N];                                                                              \
    u128 bw = 0;                                                                                \
    for (int i = 0; i < N; i++) {                                                               \
      u128 d = (u128)t[i] - PFX##_mod[i] - (uint64_t)bw;                                        \
      s[i] = (uint64_t)d;                                                                       \
      bw = (d >> 64) & 1;                                                                       \
    }                                                                                           \
    int ge = top || !bw;                                                                        \
    for (int i = 0; i < N; i++) r[i] = ge ? s[i] : t[i];                                        \
  }                                                                                             \
  static inline void PFX##_add(uint64_t* r, const uint64_t* a, const uint64_t* b) {             \
    uint64_t t[N];                                                                              \
    u128 c = 0;                                                                                 \
    for (int i = 0; i < N; i++) {                                                               \
      c += (u128)a[i] + b[i];                                                                   \
      t[i] = (uint64_t)c;                                                                       \
      c >>= 64;                                                                                 \
    }                                                                                           \
    PFX##_csub(r, t, (uint64_t)c);                                                              \
  }                                                                                             \
  static inline void PFX##_sub(uint64_t* r, const uint64_t* a, const uint64_t* b) {             \
    uint64_t t[N];                                                                              \
    u128 bw = 0;                                                                                \
    for (int i = 0; i < N; i++) {                                                               \
      u128 d = (u128)a[i] - b[i] - (uint64_t)bw;                                                \
      t[i] = (uint64_t)d;                                                                       \
      bw = (d >> 64) & 1;                                                                       \
    }                                                                                           \
    if (bw) {                                                                                   \
      u128 c = 0;                                                                               \
      for (int i = 0; i < N; i++) {                                                             \
        c += (u128)t[i] + PFX##_mod[i];                                                         \
        t[i] = (uint64_t)c;                                                                     \
        c >>= 64;                                                                               \
      }                                                                                         \
    }                                                                                           \
    memcpy(r, t, sizeof(t));                                                                    \
  }                                                                                             \
  /* Montgomery product, "no-carry" CIOS: valid because the top bit of the modulus is clear    \
   * (the optimisation ark-ff's MontBackend applies to both BLS12-381 fields). */              \
  static inline void PFX##_mul(uint64_t* r, const uint64_t* a, const uint64_t* b) {             \
    uint64_t t[N] = {0};                                                                        \
    for (int i = 0; i < N; i++) {                                                               \
      u128 c1 = (u128)a[0] * b[i] + t[0];                                                       \
      const uint64_t m = (uint64_t)c1 * INV;                                                    \
      u128 c2 = (u128)m * PFX##_mod[0] + (uint64_t)c1;                                          \
      for (int j = 1; j < N; j++) {                                                             \
        c1 = (u128)a[j] * b[i] + t[j] + (uint64_t)(c1 >> 64);                                   \
        c2 = (u128)m * PFX##_mod[j] + (uint64_t)c1 + (uint64_t)(c2 >> 64);                      \
        t[j - 1] = (uint64_t)c2;                                                                \
      }                                                                                         \
      t[N - 1] = (uint64_t)(c1 >> 64) + (uint64_t)(c2 >> 64);                                   \
    }                                                                                           \
    PFX##_csub(r, t, 0);                                                                        \
  }

#define FQ_MOD {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL, 0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}
#define FQ_ONE {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL, 0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}
#define FR_MOD {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}
#define FR_ONE {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}
DEFINE_FIELD(fq, 6, FQ_MOD, 0x89f3fffcfffcfffdULL, FQ_ONE)
DEFINE_FIELD(fr, 4, FR_MOD, 0xfffffffeffffffffULL, FR_ONE)

/* ---------------------------------------------------------------- mulx / adcx / adox products
 * arkworks' `asm` feature (ark-ff-asm, Cargo.lock:83-85) emits the Montgomery products of fields up to 6 limbs as
 * mulx / adcx / adox code; the reference's README runs its benchmarks with `--features asm`.  When this file is
 * compiled for a host that has BMI2 + ADX (`make native`: -march=native) the same instruction mix is used here: one row
 * primitive  acc[0..N] += v[0..N-1] * x  with the low halves on the CF chain and the high halves on the OF chain,
 * interleaved CIOS (row of a * b_i, row of p * m) with the accumulator kept in registers.  The portable build
 * (x86-64-v3, no ADX) keeps the unsigned __int128 code above. */
#if defined(__ADX__) && defined(__BMI2__) && !defined(GO_NO_ADX)
#define GO_HAVE_ADX 1
#define MAC_ROW6(v, x, a0, a1, a2, a3, a4, a5, a6)                                                                    \
  do {                                                                                                                \
    uint64_t lo_, hi_;                                                                                                \
    __asm__("xorl %k[lo], %k[lo]\n\t"                                                                                 \
            "mulx 0(%[vp]), %[lo], %[hi]\n\tadcx %[lo], %[r0]\n\tadox %[hi], %[r1]\n\t"                                \
            "mulx 8(%[vp]), %[lo], %[hi]\n\tadcx %[lo], %[r1]\n\tadox %[hi], %[r2]\n\t"                                \
            "mulx 16(%[vp]), %[lo], %[hi]\n\tadcx %[lo], %[r2]\n\tadox %[hi], %[r3]\n\t"                               \
            "mulx 24(%[vp]), %[lo], %[hi]\n\tadcx %[lo], %[r3]\n\tadox %[hi], %[r4]\n\t"                               \
            "mulx 32(%[vp]), %[lo], %[hi]\n\tadcx %[lo], %[r4]\n\tadox %[hi], %[r5]\n\t"                               \
            "mulx 40(%[vp]), %[lo], %[hi]\n\tadcx %[lo], %[r5]\n\tadox %[hi], %[r6]\n\t"                               \
            "movl $0, %k[lo]\n\tadcx %[lo], %[r6]\n\t"                                                               \
            : [lo] "=&r"(lo_), [hi] "=&r"(hi_), [r0] "+r"(a0), [r1] "+r"(a1), [r2] "+r"(a2), [r3] "+r"(a3), [r4] "+r"(a4), \
              [r5] "+r"(a5), [r6] "+r"(a6)                                                                            \
            : [vp] "r"(v), "d"(x), "m"(*(const uint64_t(*)[6])(v))                                                    \
            : "cc");                                                                                                  \
  } while (0)
#define MAC_ROW4(v, x, a0, a1, a2, a3, a4)                                                                            \
  do {                                                                                                                \
    uint64_t lo_, hi_;                                                                                                \
    __asm__("xorl %k[lo], %k[lo]\n\t"                                                                                 \
            "mulx 0(%[vp]), %[lo], %[hi]\n\tadcx %[lo], %[r0]\n\tadox %[hi], %[r1]\n\t"                                \
            "mulx 8(%[vp]), %[lo], %[hi]\n\tadcx %[lo], %[r1]\n\tadox %[hi], %[r2]\n\t"                                \
            "mulx 16(%[vp]), %[lo], %[hi]\n\tadcx %[lo], %[r2]\n\tadox %[hi], %[r3]\n\t"                               \
            "mulx 24(%[vp]), %[lo], %[hi]\n\tadcx %[lo], %[r3]\n\tadox %[hi], %[r4]\n\t"                               \
            "movl $0, %k[lo]\n\tadcx %[lo], %[r4]\n\t"                                                               \
            : [lo] "=&r"(lo_), [hi] "=&r"(hi_), [r0] "+r"(a0), [r1] "+r"(a1), [r2] "+r"(a2), [r3] "+r"(a3), [r4] "+r"(a4) \
            : [vp] "r"(v), "d"(x), "m"(*(const uint64_t(*)[4])(v))                                                    \
            : "cc");                                                                                                  \
  } while (0)

static inline void fq_mul_adx(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0;
  uint64_t m;
#define FQ_STEP(i, c0, c1, c2, c3, c4, c5, c6)           \
  MAC_ROW6(a, b[i], c0, c1, c2, c3, c4, c5, c6);         \
  m = c0 * 0x89f3fffcfffcfffdULL;                        \
  MAC_ROW6(fq_mod, m, c0, c1, c2, c3, c4, c5, c6);       \
  c0 = 0; /* column done: it becomes the top word of the next step */
  FQ_STEP(0, t0, t1, t2, t3, t4, t5, t6)
  FQ_STEP(1, t1, t2, t3, t4, t5, t6, t0)
  FQ_STEP(2, t2, t3, t4, t5, t6, t0, t1)
  FQ_STEP(3, t3, t4, t5, t6, t0, t1, t2)
  FQ_STEP(4, t4, t5, t6, t0, t1, t2, t3)
  FQ_STEP(5, t5, t6, t0, t1, t2, t3, t4)
#undef FQ_STEP
  const uint64_t t[6] = {t6, t0, t1, t2, t3, t4};
  fq_csub(r, t, 0);
}
static inline void fr_mul_adx(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  uint64_t m;
#define FR_STEP(i, c0, c1, c2, c3, c4)                   \
  MAC_ROW4(a, b[i], c0, c1, c2, c3, c4);                 \
  m = c0 * 0xfffffffeffffffffULL;                        \
  MAC_ROW4(fr_mod, m, c0, c1, c2, c3, c4);               \
  c0 = 0;
  FR_STEP(0, t0, t1, t2, t3, t4)
  FR_STEP(1, t1, t2, t3, t4, t0)
  FR_STEP(2, t2, t3, t4, t0, t1)
  FR_STEP(3, t3, t4, t0, t1, t2)
#undef FR_STEP
  const uint64_t t[4] = {t4, t0, t1, t2};
  fr_csub(r, t, 0);
}
#define fq_mul fq_mul_adx
#define fr_mul fr_mul_adx
#endif
const char* go_build_kind(void) {
#ifdef GO_HAVE_ADX
  return "mulx/adcx/adox";
#else
  return "portable u128";
#endif
}

static void fq_inv(uint64_t* r, const uint64_t* a) { /* a^(q-2) */
  uint64_t e[6];
  memcpy(e, fq_mod, sizeof(e));
  e[0] -= 2;
  uint64_t acc[6];
  memcpy(acc, fq_one, sizeof(acc));
  for (int bit = 383; bit >= 0; bit--) {
    fq_mul(acc, acc, acc);
    if ((e[bit >> 6] >> (bit & 63)) & 1) fq_mul(acc, acc, a);
  }
  memcpy(r, acc, sizeof(acc));
}

/* ---------------------------------------------------------------- G1, Jacobian coordinates */
typedef struct { uint64_t x[6], y[6], z[6]; } jac_t; /* identity: z == 0 */

static void jac_set_identity(jac_t* p) {
  memcpy(p->x, fq_one, 48); memcpy(p->y, fq_one, 48); memset(p->z, 0, 48);
}
static void jac_double(jac_t* p) { /* dbl-2009-l, a = 0 */
  if (fq_is_zero(p->z)) return;
  uint64_t a[6], b[6], c[6], d[6], e[6], f[6], t[6];
  fq_mul(a, p->x, p->x);
  fq_mul(b, p->y, p->y);
  fq_mul(c, b, b);
  fq_add(t, p->x, b); fq_mul(t, t, t); fq_sub(t, t, a); fq_sub(t, t, c); fq_add(d, t, t);
  fq_add(e, a, a); fq_add(e, e, a);
  fq_mul(f, e, e);
  fq_mul(p->z, p->y, p->z); fq_add(p->z, p->z, p->z);
  fq_sub(p->x, f, d); fq_sub(p->x, p->x, d);
  fq_sub(t, d, p->x); fq_mul(t, e, t);
  fq_add(c, c, c); fq_add(c, c, c); fq_add(c, c, c);
  fq_sub(p->y, t, c);
}
/* p += (ax, ay) affine, sign < 0 => subtract; (0,0) = identity (madd-2007-bl) */
static void jac_add_affine(jac_t* p, const uint64_t* ax, const uint64_t* ay_in, int neg) {
  if (fq_is_zero(ax) && fq_is_zero(ay_in)) return;
  uint64_t ay[6];
  if (neg) { uint64_t z0[6] = {0}; fq_sub(ay, z0, ay_in); } else memcpy(ay, ay_in, 48);
  if (fq_is_zero(p->z)) { memcpy(p->x, ax, 48); memcpy(p->y, ay, 48); memcpy(p->z, fq_one, 48); return; }
  uint64_t z1z1[6], u2[6], s2[6], h[6], hh[6], i[6], j[6], r[6], v[6], t[6];
  fq_mul(z1z1, p->z, p->z);
  fq_mul(u2, ax, z1z1);
  fq_mul(s2, ay, p->z); fq_mul(s2, s2, z1z1);
  if (fq_eq(u2, p->x)) {
    if (fq_eq(s2, p->y)) { jac_double(p); return; }
    jac_set_identity(p); return;
  }
  fq_sub(h, u2, p->x);
  fq_mul(hh, h, h);
  fq_add(i, hh, hh); fq_add(i, i, i);
  fq_mul(j, h, i);
  fq_sub(r, s2, p->y); fq_add(r, r, r);
  fq_mul(v, p->x, i);
  fq_add(t, p->z, h); fq_mul(t, t, t); fq_sub(t, t, z1z1); fq_sub(p->z, t, hh);
  fq_mul(p->x, r, r); fq_sub(p->x, p->x, j); fq_sub(p->x, p->x, v); fq_sub(p->x, p->x, v);
  fq_mul(j, p->y, j); fq_add(j, j, j);
  fq_sub(t, v, p->x); fq_mul(t, r, t);
  fq_sub(p->y, t, j);
}
static void jac_add(jac_t* p, const jac_t* q) { /* add-2007-bl */
  if (fq_is_zero(q->z)) return;
  if (fq_is_zero(p->z)) { *p = *q; return; }
  uint64_t z1z1[6], z2z2[6], u1[6], u2[6], s1[6], s2[6], h[6], i[6], j[6], r[6], v[6], t[6];
  fq_mul(z1z1, p->z, p->z); fq_mul(z2z2, q->z, q->z);
  fq_mul(u1, p->x, z2z2); fq_mul(u2, q->x, z1z1);
  fq_mul(s1, p->y, q->z); fq_mul(s1, s1, z2z2);
  fq_mul(s2, q->y, p->z); fq_mul(s2, s2, z1z1);
  if (fq_eq(u1, u2)) {
    if (fq_eq(s1, s2)) { jac_double(p); return; }
    jac_set_identity(p); return;
  }
  fq_sub(h, u2, u1);
  fq_add(i, h, h); fq_mul(i, i, i);
  fq_mul(j, h, i);
  fq_sub(r, s2, s1); fq_add(r, r, r);
  fq_mul(v, u1, i);
  fq_add(t, p->z, q->z); fq_mul(t, t, t); fq_sub(t, t, z1z1); fq_sub(t, t, z2z2); fq_mul(p->z, t, h);
  fq_mul(p->x, r, r); fq_sub(p->x, p->x, j); fq_sub(p->x, p->x, v); fq_sub(p->x, p->x, v);
  fq_mul(s1, s1, j); fq_add(s1, s1, s1);
  fq_sub(t, v, p->x); fq_mul(t, r, t);
  fq_sub(p->y, t, s1);
}

/* ---------------------------------------------------------------- Pippenger (variable_base.rs) */
static unsigned ark_log2(size_t x) { unsigned r = 0; while (((size_t)1 << r) < x) r++; return r; }

typedef struct {
  const uint64_t* bases; const int32_t* digits; size_t n; int c; int window; int ndig; jac_t result;
} window_job;

static void* window_worker(void* arg) {
  window_job* job = (window_job*)arg;
  const size_t nb = (size_t)1 << job->c; /* variable_base.rs:135 allocates 1<<c buckets */
  jac_t* buckets = (jac_t*)malloc(nb * sizeof(jac_t));
  for (size_t b = 0; b < nb; b++) jac_set_identity(&buckets[b]);
  for (size_t i = 0; i < job->n; i++) {
    int32_t d = job->digits[i * job->ndig + job->window];
    const uint64_t* base = job->bases + 12 * i;
    if (d > 0) jac_add_affine(&buckets[d - 1], base, base + 6, 0);
    else if (d < 0) jac_add_affine(&buckets[-d - 1], base, base + 6, 1);
  }
  jac_t running, res;
  jac_set_identity(&running); jac_set_identity(&res);
  for (size_t b = nb; b-- > 0;) { jac_add(&running, &buckets[b]); jac_add(&res, &running); }
  free(buckets);
  job->result = res;
  return NULL;
}

typedef struct { window_job* jobs; int first, last; } thread_arg;
static void* thread_main(void* a) {
  thread_arg* t = (thread_arg*)a;
  for (int w = t->first; w < t->last; w++) window_worker(&t->jobs[w]);
  return NULL;
}

/* out_xy: affine result, Montgomery; all-zero = identity.  returns the window size used. */
int go_msm_g1(const uint64_t* bases, const uint64_t* scalars, size_t n, int scalars_are_bigint, int nthreads,
              uint64_t out_xy[12]) {
  memset(out_xy, 0, 96);
  if (n == 0) return 0;
  const int c = n < 32 ? 3 : (int)(ark_log2(n) * 69 / 100) + 2; /* variable_base.rs:16-19,105-109 */
  const int num_bits = 255;
  const int ndig = (num_bits + c - 1) / c;
  int32_t* digits = (int32_t*)malloc(n * (size_t)ndig * sizeof(int32_t));
  static const uint64_t one_raw[4] = {1, 0, 0, 0};
  for (size_t i = 0; i < n; i++) { /* into_bigint + make_digits (variable_base.rs:21-61) */
    uint64_t s[5];
    if (scalars_are_bigint) memcpy(s, scalars + 4 * i, 32); else fr_mul(s, scalars + 4 * i, one_raw);
    s[4] = 0;
    const uint64_t radix = 1ull << c, mask = radix - 1;
    uint64_t carry = 0;
    for (int k = 0; k < ndig; k++) {
      const int bit = k * c, w64 = bit >> 6, sh = bit & 63;
      uint64_t buf = s[w64] >> sh;
      if (sh + c > 64 && w64 < 3) buf |= s[w64 + 1] << (64 - sh);
      uint64_t coef = carry + (buf & mask);
      carry = (coef + radix / 2) >> c;
      digits[i * ndig + k] = (int32_t)((int64_t)coef - (int64_t)(carry << c));
    }
    digits[i * ndig + ndig - 1] += (int32_t)(carry << c);
  }
  window_job* jobs = (window_job*)calloc(ndig, sizeof(window_job));
  for (int w = 0; w < ndig; w++) { jobs[w].bases = bases; jobs[w].digits = digits; jobs[w].n = n; jobs[w].c = c; jobs[w].window = w; jobs[w].ndig = ndig; }
  if (nthreads < 1) nthreads = 1;
  if (nthreads > ndig) nthreads = ndig;
  /* rayon's work-stealing over <= ndig tasks is approximated by a static block split */
  pthread_t* th = (pthread_t*)malloc(nthreads * sizeof(pthread_t));
  thread_arg* ta = (thread_arg*)malloc(nthreads * sizeof(thread_arg));
  for (int t = 0; t < nthreads; t++) {
    ta[t].jobs = jobs; ta[t].first = (int)((long)ndig * t / nthreads); ta[t].last = (int)((long)ndig * (t + 1) / nthreads);
    pthread_create(&th[t], NULL, thread_main, &ta[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  /* variable_base.rs:168-175: fold windows high -> low with c doublings each */
  jac_t total = jobs[ndig - 1].result;
  for (int w = ndig - 2; w >= 0; w--) {
    for (int d = 0; d < c; d++) jac_double(&total);
    jac_add(&total, &jobs[w].result);
  }
  free(th); free(ta); free(jobs); free(digits);
  if (!fq_is_zero(total.z)) {
    uint64_t zi[6], zi2[6];
    fq_inv(zi, total.z);
    fq_mul(zi2, zi, zi);
    fq_mul(out_xy, total.x, zi2);
    fq_mul(zi2, zi2, zi);
    fq_mul(out_xy + 6, total.y, zi2);
  }
  return c;
}

/* Synthetic bases of the benchmark: P_i = [first + i] G, affine Montgomery x|y, (0,0) = identity.
 * CPU twin of k_generate_points (gemini_b200/csrc/msm.cu) so that the reference arm of bench.py can
 * build the same workload without a GPU. */
void go_generate_bases(size_t n, uint64_t first, uint64_t* out_xy) {
  static const uint64_t gx[6] = {0x5cb38790fd530c16ULL, 0x7817fc679976fff5ULL, 0x154f95c7143ba1c1ULL, 0xf0ae6acdf3d0e747ULL, 0xedce6ecc21dbf440ULL, 0x120177419e0bfb75ULL};
  static const uint64_t gy[6] = {0xbaac93d50ce72271ULL, 0x8c22631a7918fd8eULL, 0xdd595f13570725ceULL, 0x51ac582950405194ULL, 0x0e1c8c3fad0059c0ULL, 0x0bbc3efc5008a26aULL};
  jac_t p;
  jac_set_identity(&p);
  for (int bit = 63; bit >= 0; bit--) {
    jac_double(&p);
    if ((first >> bit) & 1) jac_add_affine(&p, gx, gy, 0);
  }
  const size_t B = 1024; /* batch for Montgomery's inversion trick */
  jac_t* run = (jac_t*)malloc(B * sizeof(jac_t));
  uint64_t(*pref)[6] = (uint64_t(*)[6])malloc(B * 48);
  for (size_t i0 = 0; i0 < n; i0 += B) {
    const size_t m = n - i0 < B ? n - i0 : B;
    uint64_t acc[6];
    memcpy(acc, fq_one, 48);
    for (size_t k = 0; k < m; k++) {
      run[k] = p;
      memcpy(pref[k], acc, 48);
      if (!fq_is_zero(p.z)) fq_mul(acc, acc, p.z);
      jac_add_affine(&p, gx, gy, 0);
    }
    uint64_t inv[6];
    fq_inv(inv, acc);
    for (size_t k = m; k-- > 0;) {
      uint64_t* o = out_xy + 12 * (i0 + k);
      if (fq_is_zero(run[k].z)) { memset(o, 0, 96); continue; }
      uint64_t zi[6], zi2[6];
      fq_mul(zi, inv, pref[k]);
      fq_mul(inv, inv, run[k].z);
      fq_mul(zi2, zi, zi);
      fq_mul(o, run[k].x, zi2);
      fq_mul(zi2, zi2, zi);
      fq_mul(o + 6, run[k].y, zi2);
    }
  }
  free(run); free(pref);
}

/* ---------------------------------------------------------------- Fr folds / sumcheck */
void go_fr_fold(const uint64_t* f, size_t n, const uint64_t r[4], uint64_t* out) { /* misc.rs:52-56 */
  for (size_t i = 0; i < (n + 1) / 2; i++) {
    uint64_t t[4];
    if (2 * i + 1 < n) { fr_mul(t, r, f + 4 * (2 * i + 1)); fr_add(out + 4 * i, f + 4 * (2 * i), t); }
    else memcpy(out + 4 * i, f + 4 * (2 * i), 32);
  }
}

/* full TimeProver run driven by a fixed challenge list; returns the number of messages written.
 * out_msgs: rounds x (a | b); out_final: f[0] | g[0]. */
size_t go_sumcheck_time(const uint64_t* f_in, size_t nf, const uint64_t* g_in, size_t ng, const uint64_t twist_in[4],
                        const uint64_t* challenges, size_t nch, uint64_t* out_msgs, uint64_t out_final[8]) {
  uint64_t* f = (uint64_t*)malloc((nf ? nf : 1) * 32);
  uint64_t* g = (uint64_t*)malloc((ng ? ng : 1) * 32);
  memcpy(f, f_in, nf * 32); memcpy(g, g_in, ng * 32);
  uint64_t twist[4];
  memcpy(twist, twist_in, 32);
  const size_t mx = nf > ng ? nf : ng;
  const size_t tot_rounds = ark_log2(mx); /* time_prover.rs:35-38 */
  static const uint64_t zero[4] = {0, 0, 0, 0};
  size_t round = 0;
  for (;;) {
    if (round > 0) { /* fold with the previous challenge, time_prover.rs:75-80 */
      if (round - 1 >= nch) break;
      const uint64_t* r = challenges + 4 * (round - 1);
      uint64_t rt[4];
      fr_mul(rt, r, twist);
      go_fr_fold(f, nf, rt, f); nf = (nf + 1) / 2;
      go_fr_fold(g, ng, r, g); ng = (ng + 1) / 2;
      fr_mul(twist, twist, twist);
    }
    if (round == tot_rounds) break;
    uint64_t a[4] = {0}, b[4] = {0}, tw2[4], runner[4], t1[4], t2[4];
    fr_mul(tw2, twist, twist);
    memcpy(runner, fr_one, 32);
    const size_t pf = (nf + 1) / 2, pg = (ng + 1) / 2, np = pf < pg ? pf : pg;
    for (size_t i = 0; i < np; i++) { /* time_prover.rs:105-118 */
      const uint64_t* fe = f + 8 * i; const uint64_t* ge = g + 8 * i;
      const uint64_t* fo = (2 * i + 1 < nf) ? f + 8 * i + 4 : zero;
      const uint64_t* go = (2 * i + 1 < ng) ? g + 8 * i + 4 : zero;
      fr_mul(t1, fe, ge); fr_mul(t1, t1, runner); fr_add(a, a, t1);
      fr_mul(t1, fe, go); fr_mul(t2, ge, fo); fr_mul(t2, t2, twist); fr_add(t1, t1, t2); fr_mul(t1, t1, runner); fr_add(b, b, t1);
      fr_mul(runner, runner, tw2);
    }
    memcpy(out_msgs + 8 * round, a, 32); memcpy(out_msgs + 8 * round + 4, b, 32);
    round++;
  }
  memset(out_final, 0, 64);
  if (nf) memcpy(out_final, f, 32);
  if (ng) memcpy(out_final + 4, g, 32);
  free(f); free(g);
  return round;
}
