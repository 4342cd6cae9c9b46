/*
 * mtsv_oracle.h — C API of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's
 * (FofanovLab/mtsv_tools v2.1.0) mtsv-binner read-assignment path.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product (mtsv_tools_b200/) never does.
 *
 * Parity status: the FM-index arithmetic lives in crate `bio` 3.0.0 and the file
 * format in `bincode` 1.3.3, neither of which is vendored under /root/reference
 * and no Rust toolchain exists in this image, so for backward search / locate /
 * .index bytes this oracle is a restatement of the published algorithm:
 * **parity unpinned** at that boundary.  Pinned pieces: edit distance
 * (src/align.rs:100-170 KATs), candidate windows (src/index.rs:794-857),
 * results format (src/binner.rs:439-472), and the Smith-Waterman score, which is
 * the reference's own ssw.c compiled into oracle/_ref/libssw_ref.so.
 */
#ifndef MTSV_ORACLE_H
#define MTSV_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_index orc_index;

/* Mirrors `Hit` (src/index.rs:30-40). */
typedef struct {
  uint32_t tax_id;
  uint32_t gi;
  uint64_t offset;
  uint32_t edit;
  uint32_t _pad;
} orc_hit;

/* Arguments of MGIndex::matching_tax_ids (src/index.rs:258-269). */
typedef struct {
  double edit_rate;
  uint32_t seed_size;
  uint32_t seed_gap;
  double min_seed;
  uint64_t max_hits;
  uint64_t tune_max_hits;
  int64_t max_candidates;  /* -1 = None */
  int64_t max_assignments; /* -1 = None */
} orc_params;

/* Work counters (SURVEY §8d "algorithmic work per read"). */
typedef struct {
  uint64_t seeds_searched;
  uint64_t bs_steps;      /* backward-search steps executed (each = 2 rank queries) */
  uint64_t seeds_used;    /* seeds that contributed hits (n_seeds) */
  uint64_t rows_located;  /* SA rows located */
  uint64_t lf_steps;      /* LF steps in locate */
  uint64_t candidates;    /* candidates produced by coalesce */
  uint64_t sw_calls;      /* candidates scored with SW */
  uint64_t sw_cells;      /* L*W summed over sw_calls */
  uint64_t ed_calls;      /* edit-distance DP calls */
  uint64_t ed_cells;      /* L*W summed over ed_calls */
  uint64_t window_bytes;  /* reference bytes fetched for verification */
  uint64_t hits;          /* hits emitted */
} orc_counters;

/* ---- index construction / IO (src/index.rs:491-582, src/io.rs:115-132) ---- */
orc_index* orc_index_build(const uint8_t* seqs, const uint64_t* seq_off, const uint32_t* gi,
                           const uint32_t* tax_id, uint64_t n_seqs, uint32_t occ_interval,
                           uint64_t sa_sample);
orc_index* orc_index_from_parts(const uint8_t* text, uint64_t n, const uint32_t* gi,
                                const uint32_t* tax_id, const uint64_t* start, const uint64_t* end,
                                uint64_t n_bins, const uint8_t* bwt, const uint64_t* sample,
                                uint64_t n_sample, uint64_t sa_sample, uint32_t occ_interval);
int orc_index_write(const orc_index* ix, const char* path);
orc_index* orc_index_read(const char* path);
void orc_index_free(orc_index* ix);

uint64_t orc_index_len(const orc_index* ix);            /* text length incl. '$' */
const uint8_t* orc_index_text(const orc_index* ix);
const uint8_t* orc_index_bwt(const orc_index* ix);
uint64_t orc_index_nbins(const orc_index* ix);
/* out arrays of length nbins */
void orc_index_bins(const orc_index* ix, uint32_t* gi, uint32_t* tax_id, uint64_t* start,
                    uint64_t* end);
uint64_t orc_index_sa_sample_rate(const orc_index* ix);
uint64_t orc_index_sa_sample_len(const orc_index* ix);
const uint64_t* orc_index_sa_sample(const orc_index* ix);
uint32_t orc_index_occ_interval(const orc_index* ix);

/* Plain suffix array of a '$'-terminated text (test helper). */
int orc_suffix_array(const uint8_t* text, uint64_t n, uint64_t* sa_out);

/* ---- FM-index primitives (bio 3.0.0 semantics) ---- */
/* returns 2 = Complete, 1 = Partial, 0 = Absent; lower/upper half-open */
int orc_backward_search(const orc_index* ix, const uint8_t* pat, uint64_t len, uint64_t* lower,
                        uint64_t* upper, uint64_t* steps);
uint64_t orc_occ(const orc_index* ix, uint64_t r, uint8_t a);
uint64_t orc_less(const orc_index* ix, uint8_t a);
uint64_t orc_locate(const orc_index* ix, uint64_t row, uint64_t* lf_steps);

/* ---- verification kernels ---- */
uint32_t orc_min_edit_distance(const uint8_t* p, uint64_t plen, const uint8_t* t, uint64_t tlen);
/* Profile::new + align_score(.,1,1) (ssw/src/lib.rs:36-86). kind: 0 auto, 1 force _ref, 2 force restated */
int orc_ssw_score(const uint8_t* read, uint64_t rlen, const uint8_t* ref, uint64_t reflen, int kind);
int orc_ssw_ref_available(void);
void orc_set_ssw_kind(int kind); /* SW used by the hot path: 0 auto, 1 reference ssw.c, 2 restated */

/* SeedHit::candidate_indices (src/index.rs:118-153); returns 1 and fills out if Some */
int orc_candidate_indices(uint64_t site, uint64_t q_off, uint64_t bin_start, uint64_t bin_end,
                          uint64_t read_len, uint64_t k, uint64_t* start, uint64_t* end);

/* ---- the hot path ---- */
/* MGIndex::matching_tax_ids on ONE strand of an already normalised read. Caller frees *hits with orc_free. */
int orc_matching_tax_ids(const orc_index* ix, const uint8_t* seq, uint64_t len,
                         const orc_params* p, orc_hit** hits, uint64_t* n_hits, orc_counters* ctr);
/* Worker closure of run_fastx_pipeline (src/binner.rs:77-131) over a batch:
 * normalise, forward + reverse-complement, concatenated.  hit_off has n_reads+1 entries. */
int orc_bin_reads(const orc_index* ix, const uint8_t* seqs, const uint64_t* seq_off,
                  uint64_t n_reads, const orc_params* p, int threads, orc_hit** hits,
                  uint64_t** hit_off, orc_counters* ctr);

/* write_assignments (src/binner.rs:310-379). Returns bytes written (0 when no hits), -1 if buf too small */
int64_t orc_format_assignments(const char* header, const orc_hit* hits, uint64_t n_hits,
                               int long_format, char* buf, uint64_t buflen);

/* collapse merge rule for equal read ids, default mode (src/collapse.rs:597-602): min edit per taxid */
void orc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
