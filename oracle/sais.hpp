// sais.hpp — suffix array by induced sorting (Nong, Zhang & Chan 2009), written for the oracle.
//
// TEST INFRASTRUCTURE ONLY (see mtsv_oracle.h).  Stands in for bio 3.0.0
// `suffix_array::suffix_array` (called at src/index.rs:563).  The suffix array of a text
// with a unique smallest terminator is unique, so any correct construction gives the same
// array; tests/test_oracle_fm.py checks this one against a naive suffix sort.
//
// Precondition: n >= 1 and T[n-1] is strictly smaller than every other symbol.
#pragma once
#include <cstdint>
#include <vector>

namespace orc {

template <typename Idx, typename Sym>
void sais(const Sym* T, Idx* SA, Idx n, Idx K) {
  if (n == 1) {
    SA[0] = 0;
    return;
  }
  if (n == 2) {
    SA[0] = 1;
    SA[1] = 0;
    return;
  }
  std::vector<bool> isS((size_t)n);
  isS[n - 1] = true;
  for (Idx i = n - 2; i >= 0; --i)
    isS[i] = T[i] < T[i + 1] || (T[i] == T[i + 1] && isS[i + 1]);
  auto is_lms = [&](Idx i) { return i > 0 && isS[i] && !isS[i - 1]; };

  std::vector<Idx> cnt((size_t)K, 0), bkt((size_t)K);
  for (Idx i = 0; i < n; ++i) ++cnt[(size_t)T[i]];
  auto bucket_starts = [&]() {
    Idx sum = 0;
    for (Idx c = 0; c < K; ++c) {
      bkt[c] = sum;
      sum += cnt[c];
    }
  };
  auto bucket_ends = [&]() {
    Idx sum = 0;
    for (Idx c = 0; c < K; ++c) {
      sum += cnt[c];
      bkt[c] = sum;
    }
  };
  auto induce = [&]() {
    bucket_starts();
    for (Idx i = 0; i < n; ++i) {
      Idx j = SA[i];
      if (j > 0 && !isS[j - 1]) SA[bkt[(size_t)T[j - 1]]++] = j - 1;
    }
    bucket_ends();
    for (Idx i = n - 1; i >= 0; --i) {
      Idx j = SA[i];
      if (j > 0 && isS[j - 1]) SA[--bkt[(size_t)T[j - 1]]] = j - 1;
    }
  };

  // pass 1: sort LMS substrings
  for (Idx i = 0; i < n; ++i) SA[i] = -1;
  bucket_ends();
  for (Idx i = 1; i < n; ++i)
    if (is_lms(i)) SA[--bkt[(size_t)T[i]]] = i;
  induce();

  // compact the sorted LMS suffixes to the front, name their substrings
  Idx n1 = 0;
  for (Idx i = 0; i < n; ++i)
    if (is_lms(SA[i])) SA[n1++] = SA[i];
  for (Idx i = n1; i < n; ++i) SA[i] = -1;
  Idx name = 0, prev = -1;
  for (Idx i = 0; i < n1; ++i) {
    Idx pos = SA[i];
    bool diff = prev < 0;
    if (!diff) {
      for (Idx d = 0;; ++d) {
        if (T[pos + d] != T[prev + d] || isS[pos + d] != isS[prev + d]) {
          diff = true;
          break;
        }
        if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) break;
      }
    }
    if (diff) {
      ++name;
      prev = pos;
    }
    SA[n1 + pos / 2] = name - 1;
  }
  for (Idx i = n - 1, j = n - 1; i >= n1; --i)
    if (SA[i] >= 0) SA[j--] = SA[i];

  Idx* SA1 = SA;
  Idx* s1 = SA + n - n1;
  if (name < n1) {
    sais<Idx, Idx>(s1, SA1, n1, name);
  } else {
    for (Idx i = 0; i < n1; ++i) SA1[s1[i]] = i;
  }

  // pass 2: place the sorted LMS suffixes and induce the rest
  for (Idx i = 1, j = 0; i < n; ++i)
    if (is_lms(i)) s1[j++] = i;
  for (Idx i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
  for (Idx i = n1; i < n; ++i) SA[i] = -1;
  bucket_ends();
  for (Idx i = n1 - 1; i >= 0; --i) {
    Idx j = SA[i];
    SA[i] = -1;
    SA[--bkt[(size_t)T[j]]] = j;
  }
  induce();
}

}  // namespace orc
