// mtsv-binner (B200) — host driver with the reference's command line above the C ABI.
//
// Mirrors src/bin/mtsv-binner.rs:21-331 (flags, defaults, validation, exit codes, resume-by-results
// :347-411) and the host part of src/binner.rs (gz sniffing :21-33, FASTA/FASTQ records via their
// id = header up to the first whitespace, --read-offset skipping :169-199, write_assignments :310-379).
// The per-read work (normalisation, both strands, matching_tax_ids) is one mtsvgpu_bin_batch call
// per batch of reads.  The reference's host is Rust; no Rust toolchain exists in this image, so the
// driver is C++ over the same C ABI a Rust build.rs would link.
//
// Exit codes: 0 ok, 2 query error, 3 no results path, 4 resume error, 11 write error,
// 12 read-parse error.
#include <ctype.h>
#include <errno.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <string>
#include <tuple>
#include <unordered_set>
#include <vector>

#include "../../include/mtsv_b200.h"

namespace {

bool g_verbose = false;
void logf(const char* level, const char* fmt, ...) {
  if (!g_verbose && strcmp(level, "DEBUG") == 0) return;
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "[%s mtsv_binner] ", level);
  vfprintf(stderr, fmt, ap);
  fputc('\n', stderr);
  va_end(ap);
}

// ---- FASTA / FASTQ records (bio::io::{fasta,fastq}); gz or plain via zlib (src/binner.rs:21-33) ----
struct Record {
  std::string id, seq;
};

class FastxReader {
 public:
  FastxReader(const char* path, bool fastq) : fastq_(fastq) {
    gz_ = gzopen(path, "rb");  // transparently reads uncompressed files too (magic 1f 8b sniffing)
    if (gz_) gzbuffer(gz_, 1 << 20);
  }
  ~FastxReader() {
    if (gz_) gzclose(gz_);
  }
  bool ok() const { return gz_ != nullptr; }
  // returns 1 = record, 0 = EOF, -1 = parse error
  int next(Record* r) {
    r->id.clear();
    r->seq.clear();
    if (!have_line_ && !read_line()) return 0;
    while (line_.empty()) {
      if (!read_line()) return 0;
    }
    const char lead = fastq_ ? '@' : '>';
    if (line_[0] != lead) return -1;
    size_t e = 1;
    while (e < line_.size() && !isspace((unsigned char)line_[e])) ++e;
    r->id.assign(line_, 1, e - 1);
    have_line_ = false;
    if (!fastq_) {
      while (read_line()) {
        if (!line_.empty() && line_[0] == '>') break;
        r->seq += line_;
      }
      return 1;
    }
    // FASTQ: sequence lines up to '+', then as many quality characters as bases
    bool plus = false;
    while (read_line()) {
      if (!line_.empty() && line_[0] == '+') {
        plus = true;
        have_line_ = false;
        break;
      }
      r->seq += line_;
    }
    if (!plus) return -1;
    size_t q = 0;
    while (q < r->seq.size()) {
      if (!read_line()) return -1;
      q += line_.size();
      have_line_ = false;
    }
    if (q != r->seq.size()) return -1;
    return 1;  // (the empty quality line of an empty record is skipped as a blank line by the next call)
  }

 private:
  bool read_line() {
    line_.clear();
    char buf[1 << 16];
    bool got = false;
    while (gzgets(gz_, buf, sizeof buf)) {
      got = true;
      size_t n = strlen(buf);
      bool eol = n && buf[n - 1] == '\n';
      if (eol) --n;
      if (n && buf[n - 1] == '\r') --n;
      line_.append(buf, n);
      if (eol) break;
    }
    have_line_ = got;
    return got;
  }
  gzFile gz_ = nullptr;
  bool fastq_;
  std::string line_;
  bool have_line_ = false;
};

// ---- write_assignments (src/binner.rs:310-379) ----
void format_assignments(const std::string& id, const mtsvgpu_hit* hits, uint64_t n, bool long_fmt,
                        std::string* out) {
  if (n == 0) return;  // :316-318
  char tmp[96];
  out->append(id);
  out->push_back(':');
  bool first = true;
  if (long_fmt) {
    std::map<std::tuple<uint32_t, uint32_t, uint64_t>, uint32_t> best;
    for (uint64_t i = 0; i < n; ++i) {
      auto key = std::make_tuple(hits[i].tax_id, hits[i].gi, hits[i].offset);
      auto it = best.find(key);
      if (it == best.end()) best[key] = hits[i].edit;
      else if (hits[i].edit < it->second) it->second = hits[i].edit;
    }
    for (const auto& kv : best) {
      if (!first) out->push_back(',');
      first = false;
      snprintf(tmp, sizeof tmp, "%u-%u-%llu=%u", std::get<0>(kv.first), std::get<1>(kv.first),
               (unsigned long long)std::get<2>(kv.first), kv.second);
      out->append(tmp);
    }
  } else {
    std::map<uint32_t, uint32_t> best;
    for (uint64_t i = 0; i < n; ++i) {
      auto it = best.find(hits[i].tax_id);
      if (it == best.end()) best[hits[i].tax_id] = hits[i].edit;
      else if (hits[i].edit < it->second) it->second = hits[i].edit;
    }
    for (const auto& kv : best) {
      if (!first) out->push_back(',');
      first = false;
      snprintf(tmp, sizeof tmp, "%u=%u", kv.first, kv.second);
      out->append(tmp);
    }
  }
  out->push_back('\n');
}

// ---- resume (src/bin/mtsv-binner.rs:347-411) ----
bool read_ids_from_results(const char* path, std::unordered_set<std::string>* ids, std::string* err) {
  FILE* f = fopen(path, "r");
  if (!f) {
    *err = strerror(errno);
    return false;
  }
  char* line = nullptr;
  size_t cap = 0;
  ssize_t n;
  while ((n = getline(&line, &cap, f)) >= 0) {
    std::string s(line, (size_t)n);
    while (!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
    size_t b = 0;
    while (b < s.size() && isspace((unsigned char)s[b])) ++b;
    if (b == s.size()) continue;
    size_t colon = s.rfind(':');  // rsplitn(2, ':')
    if (colon == std::string::npos || colon == 0) {
      *err = "Missing read id";
      free(line);
      fclose(f);
      return false;
    }
    ids->insert(s.substr(0, colon));
  }
  free(line);
  fclose(f);
  return true;
}

bool resume_offset(const char* results, const char* input, bool fastq, uint64_t* offset, std::string* err) {
  struct stat sb;
  if (stat(results, &sb) != 0) {
    *offset = 0;
    return true;
  }
  std::unordered_set<std::string> ids;
  if (!read_ids_from_results(results, &ids, err)) return false;
  FastxReader rd(input, fastq);
  if (!rd.ok()) {
    *err = "cannot open input";
    return false;
  }
  Record r;
  uint64_t idx = 0, last = 0;
  bool any = false;
  int rc;
  while ((rc = rd.next(&r)) == 1) {
    if (ids.count(r.id)) {
      last = idx;
      any = true;
    }
    ++idx;
  }
  if (rc < 0) {
    *err = "parse error while scanning input";
    return false;
  }
  *offset = any ? last + 1 : 0;
  return true;
}

void usage() {
  fprintf(stderr,
          "mtsv-binner (B200)\n"
          "USAGE: mtsv-binner (--fasta <FASTA> | --fastq <FASTQ>) --index <INDEX> --results <RESULTS_PATH> [FLAGS]\n"
          "  -i, --index <INDEX>            Path to MG-index file.\n"
          "  -m, --results <RESULTS_PATH>   Path to write results file.\n"
          "      --fasta / --fastq <PATH>   Path to FASTA / FASTQ reads (gz detected automatically).\n"
          "  -e, --edit-rate <f>            [default: 0.13]\n"
          "      --seed-size <n>            [default: 18]\n"
          "      --seed-interval <n>        [default: 15]\n"
          "      --min-seed <f>             [default: 0.015]\n"
          "      --max-hits <n>             [default: 2000]\n"
          "      --tune-max-hits <n>        [default: 200]\n"
          "      --max-assignments <n>      --max-candidates <n>\n"
          "      --read-offset <n>          [default: 0]\n"
          "      --output-format default|long\n"
          "      --force-overwrite          -v\n"
          "  -t, --threads <n>              accepted for compatibility (the GPU replaces the worker pool)\n"
          "      --gpu <id>                 CUDA device [default: 0]\n"
          "      --batch-reads <n>          reads per mtsvgpu_bin_batch call [default: 4194304]\n");
}

}  // namespace

int main(int argc, char** argv) {
  const char *fasta = nullptr, *fastq = nullptr, *index = nullptr, *results = nullptr;
  mtsvgpu_params p{0.13, 18, 15, 0.015, 2000, 200, -1, -1, 2, 0};
  uint64_t read_offset = 0, batch_reads = 4u << 20;
  bool long_fmt = false, force = false, dump_reads = false;
  int device = 0;
  auto need = [&](int& i) -> const char* {
    if (i + 1 >= argc) {
      fprintf(stderr, "error: %s requires a value\n", argv[i]);
      exit(1);
    }
    return argv[++i];
  };
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--fasta" || a == "-fa" || a == "-f") fasta = need(i);
    else if (a == "--fastq" || a == "-fq") fastq = need(i);
    else if (a == "--index" || a == "-i") index = need(i);
    else if (a == "--results" || a == "-m") results = need(i);
    else if (a == "--threads" || a == "-t") (void)need(i);
    else if (a == "--edit-rate" || a == "-e") p.edit_rate = atof(need(i));
    else if (a == "--seed-size") p.seed_size = (uint32_t)strtoul(need(i), nullptr, 10);
    else if (a == "--seed-interval") p.seed_gap = (uint32_t)strtoul(need(i), nullptr, 10);
    else if (a == "--min-seed") p.min_seed = atof(need(i));
    else if (a == "--max-hits") p.max_hits = strtoull(need(i), nullptr, 10);
    else if (a == "--tune-max-hits") p.tune_max_hits = strtoull(need(i), nullptr, 10);
    else if (a == "--max-assignments") p.max_assignments = strtoll(need(i), nullptr, 10);
    else if (a == "--max-candidates") p.max_candidates = strtoll(need(i), nullptr, 10);
    else if (a == "--read-offset") read_offset = strtoull(need(i), nullptr, 10);
    else if (a == "--output-format") long_fmt = std::string(need(i)) == "long";
    else if (a == "--force-overwrite") force = true;
    else if (a == "-v") g_verbose = true;
    else if (a == "--gpu") device = atoi(need(i));
    else if (a == "--dump-reads") dump_reads = true;  // test hook: parse the input, print "id<TAB>seq", no GPU
    else if (a == "--batch-reads") batch_reads = strtoull(need(i), nullptr, 10);
    else if (a == "-h" || a == "--help") {
      usage();
      return 0;
    } else {
      fprintf(stderr, "error: unknown argument %s\n", a.c_str());
      usage();
      return 1;
    }
  }
  if (dump_reads && (fasta || fastq)) {
    FastxReader rd(fasta ? fasta : fastq, fastq != nullptr);
    if (!rd.ok()) return 2;
    Record r;
    int rc;
    uint64_t idx = 0;
    while ((rc = rd.next(&r)) == 1)
      if (idx++ >= read_offset) printf("%s\t%s\n", r.id.c_str(), r.seq.c_str());
    return rc < 0 ? 12 : 0;
  }
  if ((!fasta && !fastq) || (fasta && fastq) || !index) {
    usage();
    return 1;
  }
  // validation as src/bin/mtsv-binner.rs:151,195 (the reference panics; exit code 101 there)
  if (p.edit_rate < 0.0 || p.edit_rate > 1.0) {
    logf("ERROR", "Edit tolerance proportion must be between 0 and 1, inclusive");
    return 101;
  }
  if (p.min_seed <= 0.0 || p.min_seed > 1.0) {
    logf("ERROR", "Min seed percent must be between 0 and 1");
    return 101;
  }
  if (p.seed_size < 16) logf("WARN", "Seed size may be small enough that it causes performance issues.");
  else if (p.seed_size > 24) logf("WARN", "Seed size may be large enough that significant results are ignored.");
  if (!results) {
    logf("ERROR", "No results path provided!");
    return 3;  // :262-265
  }
  const char* input = fasta ? fasta : fastq;
  const bool is_fastq = fastq != nullptr;
  struct stat sb;
  const bool append = !force && stat(results, &sb) == 0;
  uint64_t resume = 0;
  if (!force) {
    std::string err;
    if (!resume_offset(results, input, is_fastq, &resume, &err)) {
      logf("ERROR", "Error computing resume offset: %s", err.c_str());
      return 4;  // :288-291
    }
    if (append) logf("INFO", "Existing results detected at %s; resuming after read offset %llu", results,
                     (unsigned long long)resume);
  }
  read_offset += resume;

  FILE* out = fopen(results, append ? "a" : "w");  // src/binner.rs:54-61
  if (!out) {
    logf("ERROR", "Error running query: cannot open %s: %s", results, strerror(errno));
    return 2;
  }
  logf("INFO", "Deserializing candidate filter ...");
  mtsvgpu_index* ix = nullptr;
  if (mtsvgpu_index_open(index, device, nullptr, &ix) != 0) {
    logf("ERROR", "Error running query: %s", mtsvgpu_last_error());
    fclose(out);
    return 2;
  }
  FastxReader rd(input, is_fastq);
  if (!rd.ok()) {
    logf("ERROR", "Error running query: cannot open %s", input);
    mtsvgpu_index_close(ix);
    fclose(out);
    return 2;
  }
  logf("INFO", "Beginning queries.");
  auto t0 = std::chrono::steady_clock::now();
  std::vector<uint8_t> seqs;
  std::vector<uint64_t> offs;
  std::vector<std::string> ids;
  Record r;
  uint64_t skipped = 0, total_reads = 0, total_lines = 0;
  std::string text;
  int rc = 1;
  while (rc == 1) {
    seqs.clear();
    offs.assign(1, 0);
    ids.clear();
    while (ids.size() < batch_reads && (rc = rd.next(&r)) == 1) {
      if (skipped < read_offset) {  // .skip(read_offset), src/binner.rs:176,199
        ++skipped;
        continue;
      }
      seqs.insert(seqs.end(), r.seq.begin(), r.seq.end());
      offs.push_back(seqs.size());
      ids.push_back(r.id);
    }
    if (rc < 0) {
      logf("ERROR", "Unable to read from input file: malformed record after %llu reads",
           (unsigned long long)(total_reads + ids.size() + skipped));
      mtsvgpu_index_close(ix);
      fclose(out);
      return 12;  // src/binner.rs:81-84
    }
    if (ids.empty()) break;
    mtsvgpu_hit* hits = nullptr;
    uint64_t* hit_off = nullptr;
    if (mtsvgpu_bin_batch(ix, seqs.data(), offs.data(), ids.size(), &p, &hits, &hit_off) != 0) {
      logf("ERROR", "Error running query: %s", mtsvgpu_last_error());
      mtsvgpu_index_close(ix);
      fclose(out);
      return 2;
    }
    text.clear();
    for (size_t i = 0; i < ids.size(); ++i) {
      size_t before = text.size();
      format_assignments(ids[i], hits + hit_off[i], hit_off[i + 1] - hit_off[i], long_fmt, &text);
      if (text.size() != before) ++total_lines;
    }
    mtsvgpu_free(hits);
    mtsvgpu_free(hit_off);
    if (!text.empty() && fwrite(text.data(), 1, text.size(), out) != text.size()) {
      logf("ERROR", "Error writing to result file (%s)", strerror(errno));
      mtsvgpu_index_close(ix);
      fclose(out);
      return 11;  // src/binner.rs:136-139
    }
    total_reads += ids.size();
  }
  if (fclose(out) != 0) {
    logf("ERROR", "Error writing to result file (%s)", strerror(errno));
    mtsvgpu_index_close(ix);
    return 11;
  }
  mtsvgpu_index_close(ix);
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  logf("INFO", "All reads binned: %llu reads, %llu result lines. Took %.3f seconds.",
       (unsigned long long)total_reads, (unsigned long long)total_lines, secs);
  return 0;
}
