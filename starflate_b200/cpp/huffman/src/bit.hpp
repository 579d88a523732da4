// One bit as a value type (interface parity: reference huffman/src/bit.hpp:14-65).
#pragma once

#include <cassert>
#include <compare>
#include <ostream>

namespace starflate::huffman {

class bit {
  bool v_{};

public:
  constexpr bit() = default;
  constexpr explicit bit(bool b) : v_{b} {}
  constexpr explicit bit(int i) : v_{i != 0} { assert(i == 0 || i == 1); }
  constexpr explicit bit(char c) : v_{c == '1'} { assert(c == '0' || c == '1'); }
  constexpr explicit operator bool() const { return v_; }
  constexpr explicit operator char() const { return v_ ? '1' : '0'; }
  friend constexpr auto operator<=>(const bit&, const bit&) = default;
  friend auto operator<<(std::ostream& os, bit b) -> std::ostream& { return os << static_cast<char>(b); }
};

namespace literals {
consteval auto operator""_b(unsigned long long n) -> bit
{
  return bit{static_cast<int>(n)};
}
}  // namespace literals

}  // namespace starflate::huffman
