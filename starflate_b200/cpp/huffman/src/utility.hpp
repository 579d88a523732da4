// starflate_b200 — C++23 host mirror of the reference's huffman/ toolkit, decode-side subset.
// Interface parity with garymm/starflate huffman/src/utility.hpp:9-45 (tags, byte_array, symbol
// concept); written from scratch for this repository.
#pragma once

#include <array>
#include <concepts>
#include <cstddef>

namespace starflate::huffman {

template <class T, std::size_t N>
using c_array = T[N];

// constructor tags of huffman::table (reference: utility.hpp:17-30)
struct table_contents_tag { explicit table_contents_tag() = default; };
inline constexpr table_contents_tag table_contents{};
struct symbol_bitsize_tag { explicit symbol_bitsize_tag() = default; };
inline constexpr symbol_bitsize_tag symbol_bitsize{};

// std::array<std::byte, N> from integer-like values (reference: utility.hpp:32-36)
template <class... Ts>
constexpr auto byte_array(Ts... values)
{
  return std::array<std::byte, sizeof...(Ts)>{static_cast<std::byte>(values)...};
}

template <class T>
concept symbol = std::regular<T> && std::totally_ordered<T>;

}  // namespace starflate::huffman
