// A Huffman code word: `bitsize` bits, value stored MSB-first as read from the stream.
// Interface parity with the reference's huffman/src/code.hpp:19-118 (bitsize(), value(),
// left-append of a bit, ordering by (bitsize, value), binary-literal construction).
#pragma once

#include "huffman/src/bit.hpp"

#include <cassert>
#include <compare>
#include <cstddef>
#include <cstdint>
#include <ostream>

namespace starflate::huffman {

class code {
  std::size_t bitsize_{0};
  std::size_t value_{0};

public:
  constexpr code() = default;
  constexpr code(std::size_t bitsize, std::size_t value) : bitsize_{bitsize}, value_{value} {}

  constexpr auto bitsize() const -> std::size_t { return bitsize_; }
  constexpr auto value() const -> std::size_t { return value_; }

  /// append a bit on the right (the next bit read from the stream)
  friend constexpr auto operator<<(code c, bit b) -> code
  {
    return {c.bitsize_ + 1, (c.value_ << 1) | (static_cast<bool>(b) ? 1u : 0u)};
  }
  friend constexpr auto operator<=>(const code& a, const code& b)
  {
    if (const auto c = a.bitsize_ <=> b.bitsize_; c != 0) return c;
    return a.value_ <=> b.value_;
  }
  friend constexpr auto operator==(const code&, const code&) -> bool = default;
  friend auto operator<<(std::ostream& os, const code& c) -> std::ostream&
  {
    for (std::size_t i = c.bitsize_; i-- > 0;) os << (((c.value_ >> i) & 1u) ? '1' : '0');
    return os;
  }
};

namespace literals {
template <char... Cs>
consteval auto operator""_c() -> code
{
  code c{};
  ((c = c << bit{Cs}), ...);
  return c;
}
}  // namespace literals

}  // namespace starflate::huffman
