// mtsv_oracle.cpp — CPU restatement of mtsv-binner's read-assignment path.
//
// TEST INFRASTRUCTURE ONLY (see mtsv_oracle.h for who may load this and for the
// "parity unpinned" statement about the rust-bio / bincode boundary).
//
// Every function cites the reference lines (under /root/reference) it restates.  The data
// structures deliberately keep rust-bio's shape and cost (byte BWT, per-symbol u64
// occurrence checkpoints, row-sampled suffix array, full-matrix u32 edit-distance DP) so the
// multi-threaded batch entry point can double as the CPU baseline.
#include "mtsv_oracle.h"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "sais.hpp"

namespace {

struct Bin {  // src/index.rs:44-54
  uint32_t gi, tax_id;
  uint64_t start, end;
};

}  // namespace

struct orc_index {  // MGIndex (src/index.rs:60-68) + SampledSuffixArray (bio 3.0.0)
  std::vector<uint8_t> text;  // `sequences`, last byte '$'
  std::vector<Bin> bins;
  std::vector<uint8_t> bwt;
  std::vector<uint64_t> less;               // bio::bwt::less
  std::vector<std::vector<uint64_t>> occ;   // bio::bwt::Occ.occ  [symbol][checkpoint]
  uint32_t k = 64;                          // Occ.k
  std::vector<uint64_t> sample;             // SampledSuffixArray.sample (rows 0,s,2s..)
  uint64_t s = 32;                          // SampledSuffixArray.s
  std::unordered_map<uint64_t, uint64_t> extra_rows;
  uint8_t sentinel = '$';
};

namespace {

// ------------------------------------------------------------------ FM primitives

// bio 3.0.0 Occ::get: occurrences of `a` in bwt[0..=r]  
inline uint64_t occ_get(const orc_index& ix, uint64_t r, uint8_t a) {
  uint64_t cp = r / ix.k;
  uint64_t c = ix.occ[a][cp];
  const uint8_t* b = ix.bwt.data();
  for (uint64_t i = cp * ix.k + 1; i <= r; ++i) c += (b[i] == a);
  return c;
}

// bio 3.0.0 FMIndexable::backward_search, as called at src/index.rs:305
int backward_search(const orc_index& ix, const uint8_t* pat, uint64_t len, uint64_t* lower,
                    uint64_t* upper, uint64_t* steps) {
  uint64_t l = 0, r = ix.bwt.size() - 1;
  uint64_t pl = l, pr = r;
  uint64_t matched = 0, nsteps = 0;
  for (uint64_t i = len; i-- > 0;) {
    uint8_t a = pat[i];
    uint64_t less = ix.less[a];
    pl = l;
    pr = r;
    l = less + (l > 0 ? occ_get(ix, l - 1, a) : 0);
    r = less + occ_get(ix, r, a) - 1;  // wrapping, as release-mode usize
    ++nsteps;
    if (l == r + 1) break;
    ++matched;
  }
  if (steps) *steps = nsteps;
  if (matched == len && len > 0) {
    *lower = l;
    *upper = r + 1;
    return 2;
  }
  if (matched == 0) {
    *lower = *upper = 0;
    return 0;
  }
  *lower = pl;
  *upper = pr + 1;
  return 1;
}

// bio 3.0.0 SampledSuffixArray::get  
inline uint64_t locate(const orc_index& ix, uint64_t row, uint64_t* lf_steps) {
  uint64_t pos = row, off = 0;
  for (;;) {
    if (pos % ix.s == 0) {
      if (lf_steps) *lf_steps += off;
      return ix.sample[pos / ix.s] + off;
    }
    uint8_t c = ix.bwt[pos];
    if (c == ix.sentinel) {
      if (lf_steps) *lf_steps += off;
      return ix.extra_rows.at(pos) + off;
    }
    pos = ix.less[c] + occ_get(ix, pos - 1, c);
    ++off;
  }
}

// ------------------------------------------------------------------ verification

// Aligner::min_edit_distance (src/align.rs:28-85): full (|p|+1)x(|t|+1) u32 matrix, first row 0,
// first column i, unit costs, min of last row.
uint32_t min_edit_distance(std::vector<uint32_t>& d, const uint8_t* p, size_t plen,
                           const uint8_t* t, size_t tlen) {
  size_t row_mult = tlen + 1;
  d.resize((plen + 1) * row_mult, 0);
  for (size_t i = 0; i < row_mult; ++i) d[i] = 0;
  for (size_t row = 1; row <= plen; ++row) d[row * row_mult] = (uint32_t)row;
  for (size_t row = 1; row <= plen; ++row) {
    uint8_t pc = p[row - 1];
    const uint32_t* up = &d[(row - 1) * row_mult];
    uint32_t* cur = &d[row * row_mult];
    for (size_t col = 1; col <= tlen; ++col) {
      uint32_t delta = pc != t[col - 1];
      cur[col] = std::min(up[col - 1] + delta, std::min(up[col] + 1, cur[col - 1] + 1));
    }
  }
  const uint32_t* last = &d[plen * row_mult];
  return *std::min_element(last, last + row_mult);
}

// ---- reference ssw.c, loaded from oracle/_ref/libssw_ref.so when it was built ----
struct SswApi {
  void* (*init)(const int8_t*, int32_t, const int8_t*, int32_t, int8_t) = nullptr;
  void* (*align)(const void*, const int8_t*, int32_t, uint8_t, uint8_t, uint8_t, uint16_t,
                 int32_t, int32_t) = nullptr;
  void (*init_destroy)(void*) = nullptr;
  void (*align_destroy)(void*) = nullptr;
  bool ok = false;
};

const SswApi& ssw_api() {
  static SswApi api = [] {
    SswApi a;
    Dl_info info;
    std::string dir = ".";
    if (dladdr((void*)&orc_free, &info) && info.dli_fname) {
      std::string f = info.dli_fname;
      size_t sl = f.rfind('/');
      dir = sl == std::string::npos ? "." : f.substr(0, sl);
    }
    std::string path = dir + "/_ref/libssw_ref.so";
    void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) return a;
    a.init = (decltype(a.init))dlsym(h, "ssw_init");
    a.align = (decltype(a.align))dlsym(h, "ssw_align");
    a.init_destroy = (decltype(a.init_destroy))dlsym(h, "init_destroy");
    a.align_destroy = (decltype(a.align_destroy))dlsym(h, "align_destroy");
    a.ok = a.init && a.align && a.init_destroy && a.align_destroy;
    return a;
  }();
  return api;
}

int g_ssw_kind = 0;  // 0 auto (reference ssw.c when built), 1 force reference, 2 force restated

// ssw/src/lib.rs:11-16 (note the last entry is +1: N-N scores as a match)
const int8_t kIdentMatrix[25] = {1, -1, -1, -1, -1, -1, 1, -1, -1, -1, -1, -1, 1,
                                 -1, -1, -1, -1, -1, 1, -1, -1, -1, -1, -1, 1};

// ssw/src/lib.rs:89-105
inline int8_t dna5(uint8_t b) {
  switch (b) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return 4;
  }
}

// Textbook local Smith-Waterman under the parameters the reference passes to SSW
// (match +1 incl. N-N, mismatch -1, gap open 1 + extend 1 per ssw.c:213-220 => -1 per base).
// Equal to ssw_align(...).score1 whenever the score stays below 255 - bias.
int sw_restated(const int8_t* read, size_t rlen, const int8_t* ref, size_t reflen) {
  std::vector<int> prev(rlen + 1, 0), cur(rlen + 1, 0);
  int best = 0;
  for (size_t j = 1; j <= reflen; ++j) {
    cur[0] = 0;
    for (size_t i = 1; i <= rlen; ++i) {
      int s = read[i - 1] == ref[j - 1] ? 1 : -1;
      int h = std::max(0, prev[i - 1] + s);
      h = std::max(h, prev[i] - 1);
      h = std::max(h, cur[i - 1] - 1);
      cur[i] = h;
      best = std::max(best, h);
    }
    std::swap(prev, cur);
  }
  return best;
}

// Profile (ssw/src/lib.rs:20-86): one per read, reused across candidates.
struct Profile {
  std::vector<int8_t> read_num;
  void* raw = nullptr;
  bool use_ref;
  Profile(const uint8_t* read, size_t len, int kind) {
    read_num.resize(len);
    for (size_t i = 0; i < len; ++i) read_num[i] = dna5(read[i]);
    use_ref = kind == 1 || (kind == 0 && ssw_api().ok);
    if (use_ref) raw = ssw_api().init(read_num.data(), (int32_t)len, kIdentMatrix, 5, 2);
  }
  ~Profile() {
    if (raw) ssw_api().init_destroy(raw);
  }
  int align_score(const uint8_t* ref, size_t reflen) const {
    std::vector<int8_t> ref_num(reflen);
    for (size_t i = 0; i < reflen; ++i) ref_num[i] = dna5(ref[i]);
    if (use_ref) {
      void* a = ssw_api().align(raw, ref_num.data(), (int32_t)reflen, 1, 1, 0, 0, 0,
                                (int32_t)(read_num.size() / 2));
      int score = *(const uint16_t*)a;  // s_align.score1 (ssw/src/ssw.h)
      ssw_api().align_destroy(a);
      return score;
    }
    return sw_restated(read_num.data(), read_num.size(), ref_num.data(), reflen);
  }
};

// ------------------------------------------------------------------ candidate windows

// SeedHit::candidate_indices (src/index.rs:118-153), release-mode wrapping arithmetic
inline bool candidate_indices(uint64_t site, uint64_t seed_offset, const Bin& bin,
                              uint64_t read_len, uint64_t k, uint64_t* s, uint64_t* e) {
  uint64_t start_offset = seed_offset + k;
  uint64_t cand_start =
      ((uint64_t)(site - start_offset) < bin.start || start_offset > site) ? bin.start
                                                                           : site - start_offset;
  uint64_t cand_end = site + (read_len - seed_offset) + k;
  if (cand_end > bin.end) cand_end = bin.end;
  if (cand_start > cand_end || cand_start < bin.start || cand_end > bin.end ||
      cand_end - cand_start < read_len - k)
    return false;
  *s = cand_start;
  *e = cand_end;
  return true;
}

struct SeedHit {  // src/index.rs:109-113, derived Ord = (reference_offset, query_offset)
  uint64_t reference_offset, query_offset;
  bool operator<(const SeedHit& o) const {
    return reference_offset != o.reference_offset ? reference_offset < o.reference_offset
                                                  : query_offset < o.query_offset;
  }
};

struct Candidate {  // ReferenceCandidate (src/index.rs:158-165)
  uint64_t start, end;
  size_t bin;
  uint64_t num_seeds;
};

// coalesce_seed_sites (src/index.rs:435-487) with ReferenceCandidate::{new,add_seed_hit} (:169-236)
std::vector<Candidate> coalesce(const orc_index& ix, std::vector<SeedHit>& hits, uint64_t min_seeds,
                                uint64_t read_len, uint64_t k) {
  std::sort(hits.begin(), hits.end());
  std::vector<Candidate> out;
  bool have = false;
  Candidate cur{};
  size_t b = 0;
  for (const SeedHit& sh : hits) {
    while (ix.bins[b].end <= sh.reference_offset) ++b;  // :455-458
    uint64_t ws, we;
    bool some = candidate_indices(sh.reference_offset, sh.query_offset, ix.bins[b], read_len, k,
                                  &ws, &we);
    if (have) {
      bool merged = false;
      if (some && b == cur.bin &&
          ((cur.start <= ws && ws < cur.end) || (cur.start < we && we <= cur.end))) {  // :216-221
        cur.start = std::min(cur.start, ws);
        cur.end = std::max(cur.end, we);
        cur.num_seeds += 1;
        merged = true;
      }
      if (!merged) {
        if (cur.num_seeds >= min_seeds) out.push_back(cur);  // :467-469
        have = some;                                         // :472
        if (some) cur = Candidate{ws, we, b, 1};
      }
    } else {
      have = some;  // :475
      if (some) cur = Candidate{ws, we, b, 1};
    }
  }
  if (have && cur.num_seeds >= min_seeds) out.push_back(cur);  // :481-485
  return out;
}

// ------------------------------------------------------------------ the hot function

struct Scratch {
  std::vector<uint32_t> dp;
  std::vector<SeedHit> seed_hits;
  std::vector<uint8_t> seq_no_n;
};

// MGIndex::matching_tax_ids (src/index.rs:258-432)
void matching_tax_ids(const orc_index& ix, const uint8_t* seq, uint64_t len, const orc_params& p,
                      Scratch& sc, std::vector<orc_hit>& out, orc_counters* ctr) {
  // :272-279
  sc.seq_no_n.assign(seq, seq + len);
  for (auto& b : sc.seq_no_n)
    if (b == 'N') b = '.';
  // :281-282
  uint64_t k = (uint64_t)std::ceil((double)len * p.edit_rate);
  const uint64_t S = p.seed_size, G = p.seed_gap;

  // :284-286 — (0..len+1-S).step(G). The reference panics when len < S-1 (usize wrap then
  // out-of-range slice); this restatement defines that case as "no seeds".
  uint64_t n_starts = len + 1 >= S ? len + 1 - S : 0;
  std::vector<SeedHit>& bin_locations = sc.seed_hits;
  bin_locations.clear();
  double n_seeds = 0.0;
  uint64_t next_offset = 0, seed_interval = G;
  // (itertools `.step(0)` panics in the reference; a zero gap is treated here as one seed)
  for (uint64_t offset = 0; offset < n_starts; offset += (G ? G : n_starts)) {
    if (offset < next_offset) continue;  // :300-302
    uint64_t lo = 0, up = 0, steps = 0;
    int r = backward_search(ix, seq + offset, S, &lo, &up, &steps);  // :305
    if (ctr) {
      ctr->seeds_searched++;
      ctr->bs_steps += steps;
    }
    if (r != 2) continue;             // :312-332 (only Complete sets the bounds)
    if (lo == 0 && up == 0) continue; // :330
    uint64_t n_hits = up - lo;
    if (n_hits > p.max_hits) continue;  // :335-337
    if (n_hits > p.tune_max_hits) {     // :338-344
      seed_interval *= 2;
      next_offset = offset + seed_interval;
    }
    for (uint64_t row = lo; row < up; ++row) {  // :347-352
      uint64_t lf = 0;
      uint64_t pos = locate(ix, row, &lf);
      if (ctr) {
        ctr->rows_located++;
        ctr->lf_steps += lf;
      }
      bin_locations.push_back(SeedHit{pos, offset});
    }
    n_seeds += 1.0;  // :354
    if (ctr) ctr->seeds_used++;
  }
  // :358
  uint64_t min_seeds = (uint64_t)std::max(std::floor(n_seeds * p.min_seed), 1.0);
  // :362-366
  std::vector<Candidate> refs = coalesce(ix, bin_locations, min_seeds, len, k);
  // :369 — sort_by is stable
  std::stable_sort(refs.begin(), refs.end(),
                   [](const Candidate& a, const Candidate& b) { return a.num_seeds > b.num_seeds; });
  if (ctr) ctr->candidates += refs.size();

  std::vector<uint32_t> matches;
  size_t out_begin = out.size();
  if (len == 0) return;  // Profile::new asserts read.len() > 0 (ssw/src/lib.rs:37); no seeds anyway
  Profile* profile = nullptr;  // built lazily: construction has no observable effect
  uint64_t candidates_checked = 0;
  const uint64_t threshold = len - k * 2;  // :406 wrapping
  for (const Candidate& c : refs) {
    if (p.max_candidates >= 0 && candidates_checked >= (uint64_t)p.max_candidates) break;  // :385-389
    candidates_checked++;
    uint32_t tax = ix.bins[c.bin].tax_id;
    if (std::find(matches.begin(), matches.end(), tax) != matches.end()) continue;  // :393-396
    const uint8_t* cand_seq = ix.text.data() + c.start;  // :401
    size_t cand_len = c.end - c.start;
    if (!profile) profile = new Profile(seq, len, g_ssw_kind);
    int score = profile->align_score(cand_seq, cand_len);  // :402
    if (ctr) {
      ctr->sw_calls++;
      ctr->sw_cells += len * cand_len;
      ctr->window_bytes += cand_len;
    }
    if ((uint64_t)score >= threshold) {  // :406
      uint32_t edits = min_edit_distance(sc.dp, sc.seq_no_n.data(), len, cand_seq, cand_len);  // :409
      if (ctr) {
        ctr->ed_calls++;
        ctr->ed_cells += len * cand_len;
      }
      if ((uint64_t)edits <= k) {  // :410
        matches.push_back(tax);
        orc_hit h;
        h.tax_id = tax;
        h.gi = ix.bins[c.bin].gi;
        h.offset = c.start >= ix.bins[c.bin].start ? c.start - ix.bins[c.bin].start : 0;  // :416
        h.edit = edits;
        h._pad = 0;
        out.push_back(h);
        if (p.max_assignments >= 0 && out.size() - out_begin >= (uint64_t)p.max_assignments)
          break;  // :421-425
      }
    }
  }
  delete profile;
  if (ctr) ctr->hits += out.size() - out_begin;
}

// src/binner.rs:88-100
inline uint8_t normalise(uint8_t b) {
  switch (b) {
    case 'A': case 'a': return 'A';
    case 'C': case 'c': return 'C';
    case 'G': case 'g': return 'G';
    case 'T': case 't': return 'T';
    default: return 'N';
  }
}
// bio::alphabets::dna::revcomp restricted to the normalised alphabet (src/binner.rs:115)
inline uint8_t complement(uint8_t b) {
  switch (b) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return 'N';
  }
}

void add_counters(orc_counters* dst, const orc_counters& s) {
  dst->seeds_searched += s.seeds_searched;
  dst->bs_steps += s.bs_steps;
  dst->seeds_used += s.seeds_used;
  dst->rows_located += s.rows_located;
  dst->lf_steps += s.lf_steps;
  dst->candidates += s.candidates;
  dst->sw_calls += s.sw_calls;
  dst->sw_cells += s.sw_cells;
  dst->ed_calls += s.ed_calls;
  dst->ed_cells += s.ed_cells;
  dst->window_bytes += s.window_bytes;
  dst->hits += s.hits;
}

// ------------------------------------------------------------------ bincode 1.3.3 (legacy fixint LE)

struct Writer {
  FILE* f;
  bool ok = true;
  void raw(const void* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) ok = false;
  }
  void u8(uint8_t v) { raw(&v, 1); }
  void u32(uint32_t v) { raw(&v, 4); }
  void u64(uint64_t v) { raw(&v, 8); }
};
struct Reader {
  FILE* f;
  bool ok = true;
  void raw(void* p, size_t n) {
    if (n && fread(p, 1, n, f) != n) ok = false;
  }
  uint8_t u8() { uint8_t v = 0; raw(&v, 1); return v; }
  uint32_t u32() { uint32_t v = 0; raw(&v, 4); return v; }
  uint64_t u64() { uint64_t v = 0; raw(&v, 8); return v; }
};

}  // namespace


namespace {
// less (src/index.rs:570) and Occ::new (src/index.rs:571) from the BWT
void build_less_occ(orc_index* ix, uint32_t occ_interval) {
  const uint64_t n = ix->bwt.size();
  ix->k = occ_interval;
  // :570 less over n_alphabet() = "ACGTNacgtn": max symbol 't' (116) -> 118 entries
  const size_t m_less = 118, m_occ = 117;
  ix->less.assign(m_less, 0);
  {
    std::vector<uint64_t> count(256, 0);
    for (uint8_t c : ix->bwt) count[c]++;
    uint64_t sum = 0;
    for (size_t c = 0; c < m_less; ++c) {
      ix->less[c] = sum;
      sum += count[c];
    }
  }
  // :571 Occ::new — entry j of symbol a = count of a in bwt[0..=j*k]; alphabet symbols + '$'
  ix->occ.assign(m_occ, {});
  {
    const uint8_t alpha[] = {'A', 'C', 'G', 'T', 'N', 'a', 'c', 'g', 't', 'n', '$'};
    std::vector<uint64_t> cur(256, 0);
    for (uint64_t i = 0; i < n; ++i) {
      cur[ix->bwt[i]]++;
      if (i % occ_interval == 0)
        for (uint8_t a : alpha) ix->occ[a].push_back(cur[a]);
    }
  }
}
}  // namespace

// =================================================================== C API

extern "C" {

void orc_free(void* p) { free(p); }

int orc_suffix_array(const uint8_t* text, uint64_t n, uint64_t* sa_out) {
  if (n == 0) return -1;
  std::vector<int64_t> sa(n);
  orc::sais<int64_t, uint8_t>(text, sa.data(), (int64_t)n, 256);
  for (uint64_t i = 0; i < n; ++i) sa_out[i] = (uint64_t)sa[i];
  return 0;
}

// MGIndex::new (src/index.rs:491-582) fed by parse_fasta_db (src/io.rs:135-150): sequences are
// grouped by TaxID ascending (BTreeMap), file order within a TaxID.
orc_index* orc_index_build(const uint8_t* seqs, const uint64_t* seq_off, const uint32_t* gi,
                           const uint32_t* tax_id, uint64_t n_seqs, uint32_t occ_interval,
                           uint64_t sa_sample) {
  if (occ_interval == 0 || sa_sample == 0) return nullptr;
  orc_index* ix = new orc_index;
  std::vector<uint64_t> order(n_seqs);
  for (uint64_t i = 0; i < n_seqs; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(),
                   [&](uint64_t a, uint64_t b) { return tax_id[a] < tax_id[b]; });
  uint64_t total = seq_off[n_seqs] - seq_off[0];
  ix->text.reserve(total + 1);
  for (uint64_t o : order) {  // :497-510
    Bin b;
    b.gi = gi[o];
    b.tax_id = tax_id[o];
    b.start = ix->text.size();
    ix->text.insert(ix->text.end(), seqs + seq_off[o], seqs + seq_off[o + 1]);
    b.end = ix->text.size();
    ix->bins.push_back(b);
  }
  for (auto& b : ix->text) {  // :543-553
    switch (b) {
      case 'A': case 'C': case 'G': case 'T': case 'N': break;
      case 'a': b = 'A'; break;
      case 'c': b = 'C'; break;
      case 'g': b = 'G'; break;
      case 't': b = 'T'; break;
      default: b = 'N';
    }
  }
  ix->text.push_back('$');  // :555
  const uint64_t n = ix->text.size();

  // :563 suffix_array, :567 bwt
  ix->bwt.resize(n);
  ix->k = occ_interval;
  ix->s = sa_sample;
  ix->sentinel = '$';
  auto finish = [&](auto* sa) {
    for (uint64_t r = 0; r < n; ++r) {
      uint64_t p = (uint64_t)sa[r];
      ix->bwt[r] = p ? ix->text[p - 1] : ix->text[n - 1];
    }
    // :574 SuffixArray::sample — rows 0,s,2s..; extra rows where bwt[row] is the sentinel
    for (uint64_t r = 0; r < n; ++r) {
      if (r % sa_sample == 0) ix->sample.push_back((uint64_t)sa[r]);
      else if (ix->bwt[r] == ix->sentinel) ix->extra_rows[r] = (uint64_t)sa[r];
    }
  };
  if (n < (1ull << 31)) {
    std::vector<int32_t> sa(n);
    orc::sais<int32_t, uint8_t>(ix->text.data(), sa.data(), (int32_t)n, 256);
    finish(sa.data());
  } else {
    std::vector<int64_t> sa(n);
    orc::sais<int64_t, uint8_t>(ix->text.data(), sa.data(), (int64_t)n, 256);
    finish(sa.data());
  }
  build_less_occ(ix, occ_interval);
  return ix;
}

// An MGIndex assembled from already computed parts (text incl. '$', bins, BWT, row-sampled SA):
// used when the suffix array was built elsewhere (e.g. on the GPU for the 1 Gbp benchmark index).
orc_index* orc_index_from_parts(const uint8_t* text, uint64_t n, const uint32_t* gi,
                                const uint32_t* tax_id, const uint64_t* start, const uint64_t* end,
                                uint64_t n_bins, const uint8_t* bwt, const uint64_t* sample,
                                uint64_t n_sample, uint64_t sa_sample, uint32_t occ_interval) {
  if (occ_interval == 0 || sa_sample == 0 || n == 0) return nullptr;
  if (n_sample != (n + sa_sample - 1) / sa_sample) return nullptr;
  orc_index* ix = new orc_index;
  ix->text.assign(text, text + n);
  ix->bwt.assign(bwt, bwt + n);
  for (uint64_t i = 0; i < n_bins; ++i) ix->bins.push_back(Bin{gi[i], tax_id[i], start[i], end[i]});
  ix->sample.assign(sample, sample + n_sample);
  ix->s = sa_sample;
  ix->sentinel = '$';
  // extra_rows: the row whose BWT symbol is the sentinel holds suffix 0 (bio SuffixArray::sample)
  for (uint64_t r = 0; r < n; ++r)
    if (bwt[r] == '$' && r % sa_sample != 0) ix->extra_rows[r] = 0;
  build_less_occ(ix, occ_interval);
  return ix;
}

// write_to_file (src/io.rs:125-132): bincode::serialize_into of MGIndex. Layout: SURVEY §8(b).
int orc_index_write(const orc_index* ix, const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  Writer w{f};
  w.u64(ix->text.size());
  w.raw(ix->text.data(), ix->text.size());
  w.u64(ix->bins.size());
  for (const Bin& b : ix->bins) {
    w.u32(b.gi);
    w.u32(b.tax_id);
    w.u64(b.start);
    w.u64(b.end);
  }
  w.u64(ix->bwt.size());
  w.raw(ix->bwt.data(), ix->bwt.size());
  w.u64(ix->less.size());
  w.raw(ix->less.data(), ix->less.size() * 8);
  w.u64(ix->occ.size());
  for (const auto& v : ix->occ) {
    w.u64(v.size());
    w.raw(v.data(), v.size() * 8);
  }
  w.u32(ix->k);
  w.u64(ix->sample.size());
  w.raw(ix->sample.data(), ix->sample.size() * 8);
  w.u64(ix->s);
  w.u64(ix->extra_rows.size());
  for (const auto& kv : ix->extra_rows) {
    w.u64(kv.first);
    w.u64(kv.second);
  }
  w.u8(ix->sentinel);
  bool ok = w.ok;
  if (fclose(f) != 0) ok = false;
  return ok ? 0 : -2;
}

// from_file (src/io.rs:115-122)
orc_index* orc_index_read(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return nullptr;
  Reader r{f};
  orc_index* ix = new orc_index;
  auto fail = [&]() -> orc_index* {
    fclose(f);
    delete ix;
    return nullptr;
  };
  uint64_t n = r.u64();
  if (!r.ok || n > (1ull << 40)) return fail();
  ix->text.resize(n);
  r.raw(ix->text.data(), n);
  uint64_t nb = r.u64();
  if (!r.ok || nb > (1ull << 32)) return fail();
  ix->bins.resize(nb);
  for (auto& b : ix->bins) {
    b.gi = r.u32();
    b.tax_id = r.u32();
    b.start = r.u64();
    b.end = r.u64();
  }
  uint64_t nbwt = r.u64();
  if (!r.ok || nbwt != n) return fail();
  ix->bwt.resize(nbwt);
  r.raw(ix->bwt.data(), nbwt);
  uint64_t nl = r.u64();
  if (!r.ok || nl > 4096) return fail();
  ix->less.resize(nl);
  r.raw(ix->less.data(), nl * 8);
  uint64_t no = r.u64();
  if (!r.ok || no > 4096) return fail();
  ix->occ.resize(no);
  for (auto& v : ix->occ) {
    uint64_t len = r.u64();
    if (!r.ok || len > n + 1) return fail();
    v.resize(len);
    r.raw(v.data(), len * 8);
  }
  ix->k = r.u32();
  uint64_t ns = r.u64();
  if (!r.ok || ns > n + 1) return fail();
  ix->sample.resize(ns);
  r.raw(ix->sample.data(), ns * 8);
  ix->s = r.u64();
  uint64_t ne = r.u64();
  if (!r.ok || ne > n) return fail();
  for (uint64_t i = 0; i < ne; ++i) {
    uint64_t a = r.u64(), b = r.u64();
    ix->extra_rows[a] = b;
  }
  ix->sentinel = r.u8();
  if (!r.ok) return fail();
  uint8_t extra;
  if (fread(&extra, 1, 1, f) != 0) return fail();  // must be at EOF
  fclose(f);
  return ix;
}

void orc_index_free(orc_index* ix) { delete ix; }
uint64_t orc_index_len(const orc_index* ix) { return ix->text.size(); }
const uint8_t* orc_index_text(const orc_index* ix) { return ix->text.data(); }
const uint8_t* orc_index_bwt(const orc_index* ix) { return ix->bwt.data(); }
uint64_t orc_index_nbins(const orc_index* ix) { return ix->bins.size(); }
void orc_index_bins(const orc_index* ix, uint32_t* gi, uint32_t* tax_id, uint64_t* start,
                    uint64_t* end) {
  for (size_t i = 0; i < ix->bins.size(); ++i) {
    gi[i] = ix->bins[i].gi;
    tax_id[i] = ix->bins[i].tax_id;
    start[i] = ix->bins[i].start;
    end[i] = ix->bins[i].end;
  }
}
uint64_t orc_index_sa_sample_rate(const orc_index* ix) { return ix->s; }
uint64_t orc_index_sa_sample_len(const orc_index* ix) { return ix->sample.size(); }
const uint64_t* orc_index_sa_sample(const orc_index* ix) { return ix->sample.data(); }
uint32_t orc_index_occ_interval(const orc_index* ix) { return ix->k; }

int orc_backward_search(const orc_index* ix, const uint8_t* pat, uint64_t len, uint64_t* lower,
                        uint64_t* upper, uint64_t* steps) {
  for (uint64_t i = 0; i < len; ++i)
    if (pat[i] >= ix->less.size() || pat[i] >= ix->occ.size() || ix->occ[pat[i]].empty())
      return -1;  // the reference would index out of bounds / panic
  return backward_search(*ix, pat, len, lower, upper, steps);
}
uint64_t orc_occ(const orc_index* ix, uint64_t r, uint8_t a) { return occ_get(*ix, r, a); }
uint64_t orc_less(const orc_index* ix, uint8_t a) { return ix->less[a]; }
uint64_t orc_locate(const orc_index* ix, uint64_t row, uint64_t* lf_steps) {
  uint64_t lf = 0;
  uint64_t p = locate(*ix, row, &lf);
  if (lf_steps) *lf_steps = lf;
  return p;
}

uint32_t orc_min_edit_distance(const uint8_t* p, uint64_t plen, const uint8_t* t, uint64_t tlen) {
  std::vector<uint32_t> d;
  return min_edit_distance(d, p, plen, t, tlen);
}

int orc_ssw_ref_available(void) { return ssw_api().ok ? 1 : 0; }
void orc_set_ssw_kind(int kind) { g_ssw_kind = kind; }

int orc_ssw_score(const uint8_t* read, uint64_t rlen, const uint8_t* ref, uint64_t reflen,
                  int kind) {
  if (rlen == 0 || reflen == 0) return -1;  // asserts at ssw/src/lib.rs:37,63
  if (kind == 1 && !ssw_api().ok) return -2;
  Profile p(read, rlen, kind);
  return p.align_score(ref, reflen);
}

int orc_candidate_indices(uint64_t site, uint64_t q_off, uint64_t bin_start, uint64_t bin_end,
                          uint64_t read_len, uint64_t k, uint64_t* start, uint64_t* end) {
  Bin b{0, 0, bin_start, bin_end};
  return candidate_indices(site, q_off, b, read_len, k, start, end) ? 1 : 0;
}

int orc_matching_tax_ids(const orc_index* ix, const uint8_t* seq, uint64_t len,
                         const orc_params* p, orc_hit** hits, uint64_t* n_hits,
                         orc_counters* ctr) {
  Scratch sc;
  std::vector<orc_hit> out;
  matching_tax_ids(*ix, seq, len, *p, sc, out, ctr);
  *n_hits = out.size();
  *hits = (orc_hit*)malloc(std::max<size_t>(1, out.size()) * sizeof(orc_hit));
  if (!out.empty()) memcpy(*hits, out.data(), out.size() * sizeof(orc_hit));
  return 0;
}

int orc_bin_reads(const orc_index* ix, const uint8_t* seqs, const uint64_t* seq_off,
                  uint64_t n_reads, const orc_params* p, int threads, orc_hit** hits,
                  uint64_t** hit_off, orc_counters* ctr) {
  if (threads < 1) threads = 1;
  std::vector<std::vector<orc_hit>> per_read(n_reads);
  std::vector<orc_counters> tctr(threads);
  for (auto& c : tctr) memset(&c, 0, sizeof c);
  std::atomic<uint64_t> next{0};
  const uint64_t chunk = 64;
  auto worker = [&](int tid) {
    Scratch sc;
    std::vector<uint8_t> fwd, rc;
    for (;;) {
      uint64_t b = next.fetch_add(chunk);
      if (b >= n_reads) break;
      uint64_t e = std::min(n_reads, b + chunk);
      for (uint64_t i = b; i < e; ++i) {
        uint64_t len = seq_off[i + 1] - seq_off[i];
        fwd.resize(len);
        rc.resize(len);
        for (uint64_t j = 0; j < len; ++j) fwd[j] = normalise(seqs[seq_off[i] + j]);  // :88-100
        for (uint64_t j = 0; j < len; ++j) rc[j] = complement(fwd[len - 1 - j]);      // :115
        std::vector<orc_hit>& out = per_read[i];
        matching_tax_ids(*ix, fwd.data(), len, *p, sc, out, ctr ? &tctr[tid] : nullptr);  // :102
        matching_tax_ids(*ix, rc.data(), len, *p, sc, out, ctr ? &tctr[tid] : nullptr);   // :116
      }
    }
  };
  if (threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  uint64_t total = 0;
  uint64_t* off = (uint64_t*)malloc((n_reads + 1) * sizeof(uint64_t));
  for (uint64_t i = 0; i < n_reads; ++i) {
    off[i] = total;
    total += per_read[i].size();
  }
  off[n_reads] = total;
  orc_hit* h = (orc_hit*)malloc(std::max<uint64_t>(1, total) * sizeof(orc_hit));
  for (uint64_t i = 0; i < n_reads; ++i)
    if (!per_read[i].empty())
      memcpy(h + off[i], per_read[i].data(), per_read[i].size() * sizeof(orc_hit));
  *hits = h;
  *hit_off = off;
  if (ctr)
    for (auto& c : tctr) add_counters(ctr, c);
  return 0;
}

// write_assignments (src/binner.rs:310-379)
int64_t orc_format_assignments(const char* header, const orc_hit* hits, uint64_t n_hits,
                               int long_format, char* buf, uint64_t buflen) {
  if (n_hits == 0) return 0;  // :316-318
  std::string line = header;
  line.push_back(':');
  char tmp[96];
  if (long_format) {
    std::map<std::tuple<uint32_t, uint32_t, uint64_t>, uint32_t> best;
    for (uint64_t i = 0; i < n_hits; ++i) {
      auto key = std::make_tuple(hits[i].tax_id, hits[i].gi, hits[i].offset);
      auto it = best.find(key);
      if (it == best.end()) best[key] = hits[i].edit;
      else if (hits[i].edit < it->second) it->second = hits[i].edit;
    }
    bool first = true;
    for (const auto& kv : best) {
      if (!first) line.push_back(',');
      first = false;
      snprintf(tmp, sizeof tmp, "%u-%u-%llu=%u", std::get<0>(kv.first), std::get<1>(kv.first),
               (unsigned long long)std::get<2>(kv.first), kv.second);
      line += tmp;
    }
  } else {
    std::map<uint32_t, uint32_t> best;
    for (uint64_t i = 0; i < n_hits; ++i) {
      auto it = best.find(hits[i].tax_id);
      if (it == best.end()) best[hits[i].tax_id] = hits[i].edit;
      else if (hits[i].edit < it->second) it->second = hits[i].edit;
    }
    bool first = true;
    for (const auto& kv : best) {
      if (!first) line.push_back(',');
      first = false;
      snprintf(tmp, sizeof tmp, "%u=%u", kv.first, kv.second);
      line += tmp;
    }
  }
  line.push_back('\n');
  if (line.size() + 1 > buflen) return -1;
  memcpy(buf, line.data(), line.size());
  buf[line.size()] = 0;
  return (int64_t)line.size();
}

}  // extern "C"
