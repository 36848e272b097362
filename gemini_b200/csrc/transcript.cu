// Fiat-Shamir on the host side of the C ABI: Merlin transcripts (STROBE-128 over Keccak-f[1600]) with the
// GeminiTranscript shorthands (/root/reference/src/transcript.rs:8-34; `merlin` 3.0.0, Cargo.lock:606-608, not vendored:
// restated from the published Merlin / STROBE specifications and pinned by Merlin's own known-answer vector in
// tests/test_transcript.py), and Sumcheck::prove (/root/reference/src/subprotocols/sumcheck/proof.rs:36-66) as ONE call:
// the round loop - device message, 64-byte D2H, transcript append, challenge, next launch - never returns to the caller's
// language between rounds.  The transcript itself stays on the host, like in the reference: per round it hashes 64 bytes.
#include <string.h>

#include <new>

#include <algorithm>
#include <atomic>
#include <chrono>

#include "common.cuh"
#include "fp.cuh"
#include "fr.cuh"

namespace {

const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull, 0x0000000080000001ull,
    0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000Aull,
    0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull, 0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull,
    0x000000000000800Aull, 0x800000008000000Aull, 0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
const int KECCAK_ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};

inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

// state: 25 little-endian lanes, lane (x, y) at index x + 5 y
void keccak_f1600(uint8_t st[200]) {
  uint64_t a[5][5];
  for (int x = 0; x < 5; x++)
    for (int y = 0; y < 5; y++) memcpy(&a[x][y], st + 8 * (x + 5 * y), 8);   // little-endian host
  for (int round = 0; round < 24; round++) {
    uint64_t c[5], d[5], b[5][5];
    for (int x = 0; x < 5; x++) c[x] = a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) a[x][y] ^= d[x];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y][(2 * x + 3 * y) % 5] = rol64(a[x][y], KECCAK_ROT[x][y]);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) a[x][y] = b[x][y] ^ (~b[(x + 1) % 5][y] & b[(x + 2) % 5][y]);
    a[0][0] ^= KECCAK_RC[round];
  }
  for (int x = 0; x < 5; x++)
    for (int y = 0; y < 5; y++) memcpy(st + 8 * (x + 5 * y), &a[x][y], 8);
}

constexpr int STROBE_R = 166;
enum : uint8_t { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };

struct Strobe128 {
  uint8_t state[200];
  uint8_t pos = 0, pos_begin = 0, cur_flags = 0;

  void init(const uint8_t* protocol_label, size_t n) {
    memset(state, 0, sizeof(state));
    const uint8_t head[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memcpy(state, head, 6);
    memcpy(state + 6, "STROBEv1.0.2", 12);
    keccak_f1600(state);
    pos = pos_begin = cur_flags = 0;
    meta_ad(protocol_label, n, false);
  }
  void run_f() {
    state[pos] ^= pos_begin;
    state[pos + 1] ^= 0x04;
    state[STROBE_R + 1] ^= 0x80;
    keccak_f1600(state);
    pos = pos_begin = 0;
  }
  void absorb(const uint8_t* data, size_t n) {
    for (size_t i = 0; i < n; i++) {
      state[pos++] ^= data[i];
      if (pos == STROBE_R) run_f();
    }
  }
  void squeeze(uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
      out[i] = state[pos];
      state[pos++] = 0;
      if (pos == STROBE_R) run_f();
    }
  }
  void begin_op(uint8_t flags, bool more) {
    if (more) return;   // continuation of the current operation (same flags)
    const uint8_t old_begin = pos_begin;
    pos_begin = pos + 1;
    cur_flags = flags;
    const uint8_t hdr[2] = {old_begin, flags};
    absorb(hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && pos != 0) run_f();
  }
  void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb(d, n); }
  void ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_A, more); absorb(d, n); }
  void prf(uint8_t* out, size_t n, bool more) { begin_op(FLAG_I | FLAG_A | FLAG_C, more); squeeze(out, n); }
};

inline void le32(uint8_t out[4], size_t n) {
  out[0] = (uint8_t)n; out[1] = (uint8_t)(n >> 8); out[2] = (uint8_t)(n >> 16); out[3] = (uint8_t)(n >> 24);
}

}  // namespace

struct gm_transcript {
  Strobe128 strobe;
  void append_message(const uint8_t* label, size_t ll, const uint8_t* msg, size_t n) {
    uint8_t len[4];
    le32(len, n);
    strobe.meta_ad(label, ll, false);
    strobe.meta_ad(len, 4, true);
    strobe.ad(msg, n, false);
  }
  void challenge_bytes(const uint8_t* label, size_t ll, uint8_t* out, size_t n) {
    uint8_t len[4];
    le32(len, n);
    strobe.meta_ad(label, ll, false);
    strobe.meta_ad(len, 4, true);
    strobe.prf(out, n, false);
  }
  // ark-serialize `serialize_uncompressed` of Fr elements: 32 bytes each, little-endian CANONICAL integer
  void append_fr(const uint8_t* label, size_t ll, const gm::Fr* mont, size_t count) {
    uint8_t buf[64];
    if (count > 2) count = 2;
    for (size_t k = 0; k < count; k++) {
      const gm::Fr c = mont[k].from_mont();
      memcpy(buf + 32 * k, c.v, 32);
    }
    append_message(label, ll, buf, 32 * count);
  }
  // GeminiTranscript::get_challenge (transcript.rs:25-33): 64 PRF bytes -> Fr::from_random_bytes (first 32 bytes
  // little-endian, top bit cleared, None when >= r -> draw again); returned in Montgomery form
  gm::Fr get_challenge(const uint8_t* label, size_t ll) {
    for (;;) {
      uint8_t bytes[64];
      challenge_bytes(label, ll, bytes, 64);
      gm::Fr v;
      memcpy(v.v, bytes, 32);
      v.v[7] &= 0x7FFFFFFFu;
      bool lt = false;   // v < r ?
      for (int j = 7; j >= 0; j--) {
        const uint32_t m = gm::FrParams::mod(j);
        if (v.v[j] != m) { lt = v.v[j] < m; break; }
      }
      if (lt) return v.to_mont();
    }
  }
};

namespace gm {
bool sc_wait_message(cudaStream_t stream, ScMailbox* mb, uint32_t seq) {
  const auto t0 = std::chrono::steady_clock::now();
  unsigned spins = 0;
  while ((int32_t)(mb->msg_seq - seq) < 0) {
    if ((++spins & 0xFFFu) == 0) {
      if (cudaStreamQuery(stream) != cudaErrorNotReady) {     // nothing is running any more: the message is there, or never will be
        cudaGetLastError();
        std::atomic_thread_fence(std::memory_order_acquire);
        return (int32_t)(mb->msg_seq - seq) >= 0;
      }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) return false;
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return true;
}
}  // namespace gm

using namespace gm;

extern "C" {

int gm_transcript_new(const uint8_t* label, size_t label_len, gm_transcript** out) {
  GM_ARG(out && (label || label_len == 0), "NULL argument");
  gm_transcript* t = new (std::nothrow) gm_transcript();
  if (!t) return GM_ERR_OOM;
  t->strobe.init(reinterpret_cast<const uint8_t*>("Merlin v1.0"), 11);
  t->append_message(reinterpret_cast<const uint8_t*>("dom-sep"), 7, label, label_len);
  *out = t;
  return GM_OK;
}
int gm_transcript_clone(const gm_transcript* t, gm_transcript** out) {
  GM_ARG(t && out, "NULL argument");
  gm_transcript* c = new (std::nothrow) gm_transcript(*t);
  if (!c) return GM_ERR_OOM;
  *out = c;
  return GM_OK;
}
int gm_transcript_free(gm_transcript* t) {
  delete t;
  return GM_OK;
}
int gm_transcript_append_message(gm_transcript* t, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t len) {
  GM_ARG(t && (label || label_len == 0) && (msg || len == 0), "NULL argument");
  GM_ARG(len <= 0xFFFFFFFFull, "message longer than 2^32 - 1 bytes");
  t->append_message(label, label_len, msg, len);
  return GM_OK;
}
int gm_transcript_challenge_bytes(gm_transcript* t, const uint8_t* label, size_t label_len, uint8_t* out, size_t n) {
  GM_ARG(t && (label || label_len == 0) && (out || n == 0), "NULL argument");
  t->challenge_bytes(label, label_len, out, n);
  return GM_OK;
}
int gm_transcript_append_fr(gm_transcript* t, const uint8_t* label, size_t label_len, const uint64_t* mont, size_t count) {
  GM_ARG(t && (label || label_len == 0) && mont && count >= 1 && count <= 2, "bad argument (1 or 2 field elements)");
  Fr v[2];
  for (size_t k = 0; k < count; k++) memcpy(v[k].v, mont + 4 * k, 32);
  t->append_fr(label, label_len, v, count);
  return GM_OK;
}
int gm_transcript_get_challenge_fr(gm_transcript* t, const uint8_t* label, size_t label_len, uint64_t out_mont[4]) {
  GM_ARG(t && (label || label_len == 0) && out_mont, "NULL argument");
  const Fr c = t->get_challenge(label, label_len);
  memcpy(out_mont, c.v, 32);
  return GM_OK;
}


// Sumcheck::prove (sumcheck/proof.rs:36-66) for any prover handle.  out_msgs: rounds x (a | b), out_challenges:
// rounds x Fr, both in Montgomery limbs; *out_rounds = number of messages; out_final = the final foldings (f | g),
// which are appended to the transcript as the reference does.  One kernel per round; the message reaches the host through
// the prover's pinned mailbox (fr.cuh), the challenge goes down as a kernel argument of the next launch.
int gm_sumcheck_prove(gm_sumcheck* p, gm_transcript* t, uint64_t* out_msgs, uint64_t* out_challenges, size_t capacity, size_t* out_rounds,
                      uint64_t out_final[8]) {
  GM_ARG(p && t && out_rounds && out_final && ((out_msgs && out_challenges) || capacity == 0), "NULL argument");
  static const uint8_t L_EVAL[] = "evaluations", L_CHAL[] = "challenge", L_FINAL[] = "final-folding";
  GM_CUDA(cudaSetDevice(p->ctx->device));
  size_t k = 0;
  uint64_t msg[8], ch[4];
  int has = 0;
  GM_TRY(gm_sumcheck_next_message(p, nullptr, msg, &has));
  while (has) {
    if (k >= capacity) { set_error("gm_sumcheck_prove: more than %zu rounds", capacity); return GM_ERR_ARG; }
    Fr ab[2];
    memcpy(ab[0].v, msg, 32);
    memcpy(ab[1].v, msg + 4, 32);
    t->append_fr(L_EVAL, sizeof(L_EVAL) - 1, ab, 2);
    Fr c = t->get_challenge(L_CHAL, sizeof(L_CHAL) - 1);
    memcpy(out_msgs + 8 * k, msg, 64);
    memcpy(out_challenges + 4 * k, c.v, 32);
    k++;
    memcpy(ch, c.v, 32);
    GM_TRY(gm_sumcheck_next_message(p, ch, msg, &has));
  }
  *out_rounds = k;
  int has_final = 0;
  GM_TRY(gm_sumcheck_final_foldings(p, out_final, &has_final));
  if (!has_final) { set_error("gm_sumcheck_prove: no final foldings (the reference unwraps a None here)"); return GM_ERR_STATE; }
  Fr ff[2];
  memcpy(ff[0].v, out_final, 32);
  memcpy(ff[1].v, out_final + 4, 32);
  t->append_fr(L_FINAL, sizeof(L_FINAL) - 1, &ff[0], 1);
  t->append_fr(L_FINAL, sizeof(L_FINAL) - 1, &ff[1], 1);
  return GM_OK;
}

}  // extern "C"
