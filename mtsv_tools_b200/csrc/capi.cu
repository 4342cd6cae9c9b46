// capi.cu — the extern "C" surface declared in include/mtsv_b200.h.
#include <stdlib.h>
#include <string.h>

#include "ctx.h"

namespace mtsv {
const char* last_error_cstr();
}

using namespace mtsv;

extern "C" {

const char* mtsvgpu_version(void) { return "mtsv_b200 0.1.0 (sm_100a)"; }
const char* mtsvgpu_last_error(void) { return last_error_cstr(); }
uint64_t mtsvgpu_launch_count(void) { return g_launches.load(); }
void mtsvgpu_free(void* p) { free(p); }

int mtsvgpu_index_open(const char* index_path, int device, const mtsvgpu_index_opts* opts,
                       mtsvgpu_index** out) {
  return index_open_file(index_path, device, opts, out);
}

int mtsvgpu_index_from_parts(const uint8_t* text, uint64_t n, const mtsvgpu_bin* bins, uint64_t n_bins,
                             const uint8_t* bwt, const uint64_t* sa_sample, uint64_t sa_sample_len,
                             uint64_t sa_rate, int device, const mtsvgpu_index_opts* opts,
                             mtsvgpu_index** out) {
  return index_from_host_parts(text, n, bins, n_bins, bwt, sa_sample, sa_sample_len, sa_rate, device, opts,
                               out);
}

void mtsvgpu_index_close(mtsvgpu_index* ix) { index_destroy(ix); }

int mtsvgpu_index_get_info(const mtsvgpu_index* ix, mtsvgpu_index_info* info) {
  if (!ix || !info) return set_error(MTSVGPU_EINVAL, "null argument");
  info->text_len = ix->ix.n;
  info->n_bins = ix->ix.n_bins;
  info->file_sa_rate = ix->ix.file_sa_rate;
  info->device_sa_rate = ix->ix.sa_rate;
  info->ktab_k = ix->ix.ktab_k;
  info->device_bytes = ix->ix.device_bytes;
  info->dollar_row = ix->ix.dollar_row;
  info->load_seconds = ix->ix.load_seconds;
  info->relayout_seconds = ix->ix.relayout_seconds;
  return 0;
}

int mtsvgpu_set_stream(mtsvgpu_index* ix, void* cuda_stream) {
  if (!ix) return set_error(MTSVGPU_EINVAL, "index is NULL");
  ix->stream = cuda_stream ? (cudaStream_t)cuda_stream : ix->own_stream;
  return 0;
}

int mtsvgpu_set_profiling(mtsvgpu_index* ix, int on) {
  if (!ix) return set_error(MTSVGPU_EINVAL, "index is NULL");
  ix->profiling = on != 0;
  return 0;
}

int mtsvgpu_last_batch_stats(const mtsvgpu_index* ix, mtsvgpu_batch_stats* stats) {
  if (!ix || !stats) return set_error(MTSVGPU_EINVAL, "null argument");
  *stats = ix->stats;
  return 0;
}

int mtsvgpu_bin_batch_device(mtsvgpu_index* ix, const uint8_t* d_seqs, const uint64_t* d_seq_off,
                             uint64_t n_reads, const mtsvgpu_params* params, const mtsvgpu_hit** d_hits,
                             const uint64_t** d_hit_off, uint64_t* n_hits) {
  return bin_batch_device(ix, d_seqs, d_seq_off, n_reads, nullptr, params, d_hits, d_hit_off, n_hits);
}

static int ensure_pinned(void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap && *p) return 0;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *cap = 0;
  size_t want = bytes + bytes / 4 + 4096;
  cudaError_t e = cudaMallocHost(p, want);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    *p = nullptr;
    return set_error(MTSVGPU_ENOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  *cap = want;
  return 0;
}

int mtsvgpu_bin_batch(mtsvgpu_index* ix, const uint8_t* seqs, const uint64_t* seq_off, uint64_t n_reads,
                      const mtsvgpu_params* params, mtsvgpu_hit** hits, uint64_t** hit_off) {
  if (!ix || !seq_off || !hits || !hit_off) return set_error(MTSVGPU_EINVAL, "null argument");
  *hits = nullptr;
  *hit_off = nullptr;
  MTSV_CUDA_TRY(cudaSetDevice(ix->ix.device));
  cudaStream_t st = ix->stream;
  const uint64_t base = seq_off[0];
  const uint64_t bytes = seq_off[n_reads] - base;
  if (bytes && !seqs) return set_error(MTSVGPU_EINVAL, "seqs is NULL");
  BatchWorkspace& ws = ix->ws;
  // ---- H2D: read bytes and offsets (rebased to 0) ----
  MTSV_TRY(ws.d_seqs.reserve(bytes + 16));
  MTSV_TRY(ws.d_seq_off.reserve((n_reads + 1) * 8));
  std::vector<uint64_t> rebased;
  const uint64_t* offs = seq_off;
  if (base != 0) {
    rebased.resize(n_reads + 1);
    for (uint64_t i = 0; i <= n_reads; ++i) rebased[i] = seq_off[i] - base;
    offs = rebased.data();
  }
  if (bytes) MTSV_CUDA_TRY(cudaMemcpyAsync(ws.d_seqs.p, seqs + base, bytes, cudaMemcpyHostToDevice, st));
  MTSV_CUDA_TRY(cudaMemcpyAsync(ws.d_seq_off.p, offs, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
  // ---- compute ----
  const mtsvgpu_hit* d_hits = nullptr;
  const uint64_t* d_hit_off = nullptr;
  uint64_t n_hits = 0;
  MTSV_TRY(bin_batch_device(ix, ws.d_seqs.as<uint8_t>(), ws.d_seq_off.as<uint64_t>(), n_reads, offs, params,
                            &d_hits, &d_hit_off, &n_hits));
  // ---- D2H ----
  mtsvgpu_hit* h_hits = (mtsvgpu_hit*)malloc((n_hits ? n_hits : 1) * sizeof(mtsvgpu_hit));
  uint64_t* h_off = (uint64_t*)malloc((n_reads + 1) * sizeof(uint64_t));
  if (!h_hits || !h_off) {
    free(h_hits);
    free(h_off);
    return set_error(MTSVGPU_ENOMEM, "host allocation of the result failed");
  }
  cudaError_t e = cudaSuccess;
  if (n_hits) e = cudaMemcpyAsync(h_hits, d_hits, n_hits * sizeof(mtsvgpu_hit), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(h_off, d_hit_off, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    free(h_hits);
    free(h_off);
    return set_error(MTSVGPU_ECUDA, "result copy failed: %s", cudaGetErrorString(e));
  }
  *hits = h_hits;
  *hit_off = h_off;
  (void)ensure_pinned;
  return 0;
}

int mtsvgpu_backward_search(mtsvgpu_index* ix, const uint8_t* pats, uint32_t pat_len, uint64_t n_pats,
                            uint64_t* lower, uint64_t* upper) {
  return backward_search_batch(ix, pats, pat_len, n_pats, lower, upper);
}

int mtsvgpu_locate(mtsvgpu_index* ix, const uint64_t* rows, uint64_t n_rows, uint64_t* pos) {
  return locate_batch(ix, rows, n_rows, pos);
}

int mtsvgpu_edit_distance(int device, const uint8_t* pats, const uint64_t* pat_off, const uint8_t* texts,
                          const uint64_t* text_off, uint64_t n_pairs, uint32_t* edits) {
  return edit_distance_batch(device, pats, pat_off, texts, text_off, n_pairs, edits);
}

}  // extern "C"
