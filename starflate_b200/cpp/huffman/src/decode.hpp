// Bit-serial decode of one symbol (interface parity: reference huffman/src/decode.hpp:39-102).
// This is the host-side, definitional form; the production decoder is the LUT / canonical
// first-count decoder in csrc/deflate_lane.cuh, which returns the same (symbol, size) for every
// input (tests/ pin both against the reference).
#pragma once

#include "huffman/src/bit_span.hpp"
#include "huffman/src/table.hpp"

#include <cstdint>

namespace starflate::huffman {

template <symbol Symbol>
class decode_result {
  Symbol symbol_{};
  std::uint8_t encoded_size_{0};

public:
  static constexpr std::uint8_t kInvalidEncodedSize = 0;
  constexpr decode_result() = default;
  constexpr decode_result(Symbol s, std::uint8_t encoded_size) : symbol_{s}, encoded_size_{encoded_size} {}
  constexpr auto has_value() const -> bool { return encoded_size_ != kInvalidEncodedSize; }
  constexpr auto symbol() const -> Symbol { return symbol_; }
  constexpr auto encoded_size() const -> std::uint8_t { return encoded_size_; }
};

/// Decode one symbol from the front of `bits` without consuming it.  encoded_size() == 0 when
/// no code matches: unassigned code, code longer than the table's longest, or input exhausted.
template <symbol Symbol, std::size_t Extent>
constexpr auto decode_one(const table<Symbol, Extent>& t, bit_span bits) -> decode_result<Symbol>
{
  code current{};
  auto pos = t.begin();
  for (const bit b : bits) {
    current = current << b;
    const auto found = t.find(current, pos);
    if (found) return {(*found)->symbol, static_cast<std::uint8_t>(current.bitsize())};
    if (found.error() == t.end()) break;
    pos = found.error();
  }
  return {};
}

}  // namespace starflate::huffman
