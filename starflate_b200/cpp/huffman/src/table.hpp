// Canonical Huffman code table built from (symbol range, bitsize) pairs.
//
// Interface parity with the decode-side subset of the reference's huffman/src/table.hpp
// (symbol_bitsize constructor :360-376, canonical assignment :177-216, table_contents
// constructor, begin/end, find :420-452); the encoder-side constructors (frequencies / data) are
// out of scope.  Storage is a flat array sorted by (bitsize, symbol) — the
// order the reference's canonicalize() produces; the CUDA table builder keeps the same order
// in its scratch as per-bitsize first-code / count / offset (csrc/deflate_lane.cuh).
#pragma once

#include "huffman/src/code.hpp"
#include "huffman/src/symbol_span.hpp"
#include "huffman/src/utility.hpp"

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <expected>
#include <ranges>
#include <span>
#include <type_traits>
#include <utility>
#include <vector>

namespace starflate::huffman {

template <symbol Symbol>
struct encoding {
  code code_{};
  Symbol symbol{};
  constexpr auto bitsize() const -> std::size_t { return code_.bitsize(); }
  constexpr auto value() const -> std::size_t { return code_.value(); }
  constexpr operator code() const { return code_; }  // NOLINT(google-explicit-constructor)
  friend constexpr auto operator==(const encoding&, const encoding&) -> bool = default;
};

namespace detail {
// fixed-capacity storage so that tables with a static Extent can be constexpr objects
template <class T, std::size_t N>
class fixed_vec {
  std::array<T, N> a_{};
  std::size_t n_{0};

public:
  using const_iterator = const T*;
  constexpr void push_back(const T& v) { a_[n_++] = v; }
  constexpr auto begin() -> T* { return a_.data(); }
  constexpr auto end() -> T* { return a_.data() + n_; }
  constexpr auto begin() const -> const T* { return a_.data(); }
  constexpr auto end() const -> const T* { return a_.data() + n_; }
  constexpr auto size() const -> std::size_t { return n_; }
};
template <class T, std::size_t Extent>
using table_storage =
    std::conditional_t<Extent == std::dynamic_extent, std::vector<T>, fixed_vec<T, Extent>>;
}  // namespace detail

template <symbol Symbol, std::size_t Extent = std::dynamic_extent>
class table {
public:
  using encoding_type = encoding<Symbol>;
  using storage_type = detail::table_storage<encoding_type, Extent>;
  using const_iterator = typename storage_type::const_iterator;

private:
  storage_type nodes_{};  // sorted by (bitsize, symbol) once canonicalized

  constexpr void canonicalize()
  {
    std::ranges::sort(nodes_.begin(), nodes_.end(), [](const encoding_type& a, const encoding_type& b) {
      return a.bitsize() != b.bitsize() ? a.bitsize() < b.bitsize() : a.symbol < b.symbol;
    });
    // RFC 1951 §3.2.2; for over-subscribed sets the arithmetic simply continues and the
    // overflowed codes (value >= 2^bitsize) can never equal a prefix read from a stream
    std::size_t next = 0, prev_size = 0;
    for (auto it = nodes_.begin(); it != nodes_.end(); ++it) {
      next <<= (it->bitsize() - prev_size);
      prev_size = it->bitsize();
      it->code_ = code{it->bitsize(), next};
      ++next;
    }
  }
  template <class Pairs>
  constexpr void fill_from_bitsizes(const Pairs& pairs)
  {
    for (const auto& p : pairs) {
      const symbol_span<Symbol> s{p.first};
      for (std::size_t i = 0; i < s.size(); ++i)
        nodes_.push_back({code{static_cast<std::size_t>(p.second), 0}, static_cast<Symbol>(s.first() + i)});
    }
    canonicalize();
  }

public:
  table() = default;

  /// from (symbol range, bitsize) pairs with bitsize > 0
  template <std::ranges::input_range R>
  constexpr table(symbol_bitsize_tag, const R& pairs) { fill_from_bitsizes(pairs); }
  template <std::size_t N>
  constexpr table(symbol_bitsize_tag, const c_array<std::pair<symbol_span<Symbol>, std::uint8_t>, N>& pairs)
  {
    fill_from_bitsizes(pairs);
  }

  /// from explicit (code, symbol) contents, kept in the order given (must be sorted by bitsize)
  template <std::size_t N>
  constexpr table(table_contents_tag, const c_array<std::pair<code, Symbol>, N>& contents)
  {
    for (const auto& [c, s] : contents) nodes_.push_back({c, s});
  }

  constexpr auto begin() const -> const_iterator { return nodes_.begin(); }
  constexpr auto end() const -> const_iterator { return nodes_.end(); }
  constexpr auto size() const -> std::size_t { return nodes_.size(); }

  /// Find `c` at or after `pos`.  On a miss the error is the first element with a larger
  /// bitsize (or end()), where a bit-serial caller resumes after reading one more bit.
  [[nodiscard]] constexpr auto find(code c, const_iterator pos) const
      -> std::expected<const_iterator, const_iterator>
  {
    using R = std::expected<const_iterator, const_iterator>;
    while (pos != end() && pos->bitsize() < c.bitsize()) ++pos;
    for (; pos != end() && pos->bitsize() == c.bitsize(); ++pos)
      if (pos->value() == c.value()) return R{std::in_place, pos};
    return R{std::unexpect, pos};
  }
  [[nodiscard]] constexpr auto find(code c) const { return find(c, begin()); }
};

template <symbol Symbol, std::size_t N>
table(table_contents_tag, const c_array<std::pair<code, Symbol>, N>&) -> table<Symbol, N>;

}  // namespace starflate::huffman
