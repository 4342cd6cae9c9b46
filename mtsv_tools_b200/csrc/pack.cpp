// pack.cpp — host side of the packed-read format (core.cuh "packed reads", mtsvgpu_bin_batch_packed): what a parser
// thread does with a record's sequence instead of copying its bytes.  Normalisation as src/binner.rs:88-100
// (upper / lower case ACGT are bases, everything else is N), three bit planes per read.  AVX2 when the CPU has it
// (32 bases per step: compares + movemask give the planes directly), the SWAR encoder of core.cuh otherwise.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/mtsv_b200.h"
#include "core.cuh"

namespace {

using mtsv::packed_plane_bytes;
using mtsv::packed_record_bytes;

inline void store_bits(uint8_t* dst, uint32_t v, uint32_t nbytes) { memcpy(dst, &v, nbytes); }

void pack_one_swar(const uint8_t* seq, uint32_t L, uint8_t* rec, bool) {
  const uint32_t pb = packed_plane_bytes(L);
  uint8_t *lo = rec, *hi = rec + pb, *nn = rec + 2 * pb;
  uint32_t j = 0;
  for (; j + 8 <= L; j += 8) {
    uint64_t x;
    memcpy(&x, seq + j, 8);
    uint32_t l, h, n;
    mtsv::encode8(x, &l, &h, &n);
    lo[j >> 3] = (uint8_t)l;
    hi[j >> 3] = (uint8_t)h;
    nn[j >> 3] = (uint8_t)n;
  }
  if (j < L) {
    uint64_t x = 0;
    memcpy(&x, seq + j, L - j);
    uint32_t l, h, n;
    mtsv::encode8(x, &l, &h, &n);
    const uint32_t m = (1u << (L - j)) - 1;
    lo[j >> 3] = (uint8_t)(l & m);
    hi[j >> 3] = (uint8_t)(h & m);
    nn[j >> 3] = (uint8_t)(n & m);
  }
}

#if defined(__x86_64__)
// `safe`: at least 32 readable bytes follow seq + 32 * floor(L / 32) (true for every read but the last few of a
// buffer), so the tail is loaded in place instead of through a bounce buffer
__attribute__((target("avx2"))) inline void planes32(const uint8_t* p, uint32_t* l, uint32_t* h, uint32_t* n) {
  const __m256i u = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(p)), _mm256_set1_epi8((char)0xDF));
  const __m256i a = _mm256_cmpeq_epi8(u, _mm256_set1_epi8('A')), c = _mm256_cmpeq_epi8(u, _mm256_set1_epi8('C')),
                g = _mm256_cmpeq_epi8(u, _mm256_set1_epi8('G')), t = _mm256_cmpeq_epi8(u, _mm256_set1_epi8('T'));
  *l = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(c, t));
  *h = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(g, t));
  *n = ~(uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_or_si256(a, c), _mm256_or_si256(g, t)));
}

__attribute__((target("avx2"))) void pack_one_avx2(const uint8_t* seq, uint32_t L, uint8_t* rec, bool safe) {
  const uint32_t pb = packed_plane_bytes(L);
  constexpr uint32_t kChunks = 16;  // reads up to 512 bases: planes are collected in registers / stack words
  if (L <= 32 * kChunks) {
    uint32_t lo[kChunks], hi[kChunks], nn[kChunks];
    uint32_t j = 0, k = 0;
    for (; j + 32 <= L; j += 32, ++k) planes32(seq + j, &lo[k], &hi[k], &nn[k]);
    if (j < L) {
      const uint32_t m = (1u << (L - j)) - 1;
      if (safe) {
        planes32(seq + j, &lo[k], &hi[k], &nn[k]);
      } else {
        uint8_t tail[32] = {0};
        memcpy(tail, seq + j, L - j);
        planes32(tail, &lo[k], &hi[k], &nn[k]);
      }
      lo[k] &= m;
      hi[k] &= m;
      nn[k] &= m;
    }
    memcpy(rec, lo, pb);
    memcpy(rec + pb, hi, pb);
    memcpy(rec + 2 * pb, nn, pb);
    return;
  }
  uint8_t *lo = rec, *hi = rec + pb, *nn = rec + 2 * pb;
  uint32_t j = 0;
  for (; j + 32 <= L; j += 32) {
    uint32_t l, h, n;
    planes32(seq + j, &l, &h, &n);
    store_bits(lo + (j >> 3), l, 4);
    store_bits(hi + (j >> 3), h, 4);
    store_bits(nn + (j >> 3), n, 4);
  }
  if (j < L) {
    uint8_t tail[32] = {0};
    memcpy(tail, seq + j, L - j);
    uint32_t l, h, n;
    planes32(tail, &l, &h, &n);
    const uint32_t m = (1u << (L - j)) - 1, nb = (L - j + 7) >> 3;
    store_bits(lo + (j >> 3), l & m, nb);
    store_bits(hi + (j >> 3), h & m, nb);
    store_bits(nn + (j >> 3), n & m, nb);
  }
}
#endif

typedef void (*PackFn)(const uint8_t*, uint32_t, uint8_t*, bool);
PackFn pick() {
#if defined(__x86_64__)
  if (__builtin_cpu_supports("avx2")) return pack_one_avx2;
#endif
  return pack_one_swar;
}

}  // namespace

extern "C" {

uint64_t mtsvgpu_packed_size(const uint64_t* seq_off, uint64_t n_reads) {
  if (!seq_off) return 0;
  uint64_t total = 0;
  for (uint64_t i = 0; i < n_reads; ++i) total += 3 * ((seq_off[i + 1] - seq_off[i] + 7) >> 3);
  return total;
}

void mtsvgpu_pack_read(const uint8_t* seq, uint32_t len, uint8_t* record) {
  static const PackFn fn = pick();
  fn(seq, len, record, false);
}

int mtsvgpu_pack_reads(const uint8_t* seqs, const uint64_t* seq_off, uint64_t n_reads, uint8_t* packed,
                       uint64_t packed_cap, uint64_t* packed_bytes, int threads) {
  if (!seq_off || (n_reads && !seqs && seq_off[n_reads] != seq_off[0])) return MTSVGPU_EINVAL;
  for (uint64_t i = 0; i < n_reads; ++i)
    if (seq_off[i + 1] < seq_off[i] || seq_off[i + 1] - seq_off[i] > 0xffffffffull) return MTSVGPU_EINVAL;
  static const PackFn fn = pick();
  int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
  if ((uint64_t)T > n_reads / 4096 + 1) T = (int)(n_reads / 4096 + 1);
  // ranges of reads per thread and where their records start
  std::vector<uint64_t> first(T + 1), at(T + 1, 0);
  for (int t = 0; t <= T; ++t) first[t] = n_reads * (uint64_t)t / (uint64_t)T;
  const bool uniform = n_reads > 0 && [&] {
    const uint64_t len = seq_off[1] - seq_off[0];
    uint64_t acc = 0;
    for (uint64_t i = 0; i < n_reads; ++i) acc |= (seq_off[i + 1] - seq_off[i]) ^ len;
    return acc == 0;
  }();
  if (uniform) {
    const uint64_t rec = packed_record_bytes((uint32_t)(seq_off[1] - seq_off[0]));
    for (int t = 0; t <= T; ++t) at[t] = first[t] * rec;
  } else {
    for (int t = 0; t < T; ++t) at[t + 1] = at[t] + mtsvgpu_packed_size(seq_off + first[t], first[t + 1] - first[t]);
  }
  if (packed_bytes) *packed_bytes = at[T];
  if (at[T] > packed_cap) return MTSVGPU_ELIMIT;
  if (at[T] == 0) return 0;  // only empty reads: nothing to write
  if (!packed) return MTSVGPU_EINVAL;
  auto work = [&](int t) {
    uint8_t* out = packed + at[t];
    const uint64_t end_of_bytes = seq_off[n_reads];
    for (uint64_t r = first[t]; r < first[t + 1]; ++r) {
      const uint32_t L = (uint32_t)(seq_off[r + 1] - seq_off[r]);
      fn(seqs + seq_off[r], L, out, seq_off[r] + (L & ~31u) + 32 <= end_of_bytes);
      out += packed_record_bytes(L);
    }
  };
  if (T <= 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
  }
  return 0;
}

}  // extern "C"
