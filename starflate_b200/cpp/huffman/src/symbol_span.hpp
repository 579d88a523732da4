// Inclusive range of consecutive symbols (reference: huffman/src/symbol_span.hpp:11-55).
#pragma once

#include "huffman/src/utility.hpp"

#include <cstddef>

namespace starflate::huffman {

template <symbol Symbol>
class symbol_span {
  Symbol first_{};
  Symbol last_{};

public:
  constexpr symbol_span(Symbol first, Symbol last) : first_{first}, last_{last} {}
  constexpr explicit symbol_span(Symbol only) : symbol_span(only, only) {}
  constexpr auto first() const -> Symbol { return first_; }
  constexpr auto last() const -> Symbol { return last_; }
  constexpr auto size() const -> std::size_t { return static_cast<std::size_t>(last_ - first_) + 1; }
};

}  // namespace starflate::huffman
