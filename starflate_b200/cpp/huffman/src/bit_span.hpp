// Non-owning view of a bit sequence, LSB-first within each byte.
// Interface parity with the reference's huffman/src/bit_span.hpp:18-183 (constructors,
// begin/end, consume, consume_to_byte_boundary, pop<T>/pop_8/pop_16, byte_data); written from
// scratch.  On the device the same role is played by sfb::BitReader (csrc/deflate_lane.cuh).
#pragma once

#include "huffman/src/bit.hpp"

#include <bit>
#include <cassert>
#include <climits>
#include <concepts>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <ranges>

namespace starflate::huffman {

class bit_span : public std::ranges::view_interface<bit_span> {
  const std::byte* data_{nullptr};
  std::size_t bit_size_{0};
  std::uint8_t bit_offset_{0};  // < CHAR_BIT

public:
  class iterator {
    const bit_span* parent_{nullptr};
    std::size_t offset_{0};  // in bits, from parent_->data_

  public:
    using difference_type = std::ptrdiff_t;
    using iterator_category = std::random_access_iterator_tag;
    using iterator_concept = std::random_access_iterator_tag;
    using value_type = bit;
    using reference = bit;
    using pointer = void;

    iterator() = default;
    constexpr iterator(const bit_span& parent, std::size_t offset) : parent_{&parent}, offset_{offset} {}

    constexpr auto operator*() const -> bit
    {
      const auto b = std::to_integer<unsigned>(parent_->data_[offset_ / CHAR_BIT]);
      return bit{((b >> (offset_ % CHAR_BIT)) & 1u) != 0};
    }
    constexpr auto operator[](difference_type n) const -> bit { return *(*this + n); }
    constexpr auto operator+=(difference_type n) -> iterator&
    {
      offset_ = static_cast<std::size_t>(static_cast<difference_type>(offset_) + n);
      return *this;
    }
    constexpr auto operator-=(difference_type n) -> iterator& { return *this += -n; }
    constexpr auto operator++() -> iterator& { return *this += 1; }
    constexpr auto operator++(int) -> iterator { auto t = *this; ++*this; return t; }
    constexpr auto operator--() -> iterator& { return *this -= 1; }
    constexpr auto operator--(int) -> iterator { auto t = *this; --*this; return t; }
    friend constexpr auto operator+(iterator i, difference_type n) -> iterator { return i += n; }
    friend constexpr auto operator+(difference_type n, iterator i) -> iterator { return i += n; }
    friend constexpr auto operator-(iterator i, difference_type n) -> iterator { return i -= n; }
    friend constexpr auto operator-(const iterator& a, const iterator& b) -> difference_type
    {
      return static_cast<difference_type>(a.offset_) - static_cast<difference_type>(b.offset_);
    }
    friend constexpr auto operator==(const iterator& a, const iterator& b) -> bool { return a.offset_ == b.offset_; }
    friend constexpr auto operator<=>(const iterator& a, const iterator& b) { return a.offset_ <=> b.offset_; }
  };

  bit_span() = default;

  /// `bit_size` bits starting `bit_offset` (< 8) bits into `data[0]`
  constexpr bit_span(const std::byte* data, std::size_t bit_size, std::uint8_t bit_offset = 0)
      : data_{data}, bit_size_{bit_size}, bit_offset_{bit_offset}
  {
    assert(bit_offset < CHAR_BIT);
  }

  /// all bits of a borrowed contiguous range of bytes
  template <std::ranges::contiguous_range R>
    requires std::ranges::borrowed_range<R> && std::same_as<std::ranges::range_value_t<R>, std::byte>
  // NOLINTNEXTLINE(bugprone-forwarding-reference-overload)
  constexpr bit_span(R&& r) : bit_span(std::ranges::data(r), std::ranges::size(r) * CHAR_BIT)
  {}

  constexpr auto begin() const -> iterator { return {*this, bit_offset_}; }
  constexpr auto end() const -> iterator { return {*this, bit_offset_ + bit_size_}; }

  /// drop the first n bits (n <= size())
  constexpr auto consume(std::size_t n) & -> bit_span&
  {
    assert(n <= bit_size_);
    const std::size_t d = bit_offset_ + n;
    data_ += d / CHAR_BIT;  // NOLINT(cppcoreguidelines-pro-bounds-pointer-arithmetic)
    bit_offset_ = static_cast<std::uint8_t>(d % CHAR_BIT);
    bit_size_ -= n;
    return *this;
  }
  constexpr auto consume(std::size_t n) && -> bit_span&& { return std::move(consume(n)); }

  constexpr auto consume_to_byte_boundary() -> void
  {
    if (bit_offset_ != 0) consume(static_cast<std::size_t>(CHAR_BIT - bit_offset_));
  }

  /// read a little-endian integer at a byte boundary and drop it
  template <std::unsigned_integral T>
  constexpr auto pop() -> T
  {
    assert(bit_offset_ == 0 && bit_size_ >= sizeof(T) * CHAR_BIT);
    T v{};
    for (std::size_t i = 0; i < sizeof(T); ++i)
      v = static_cast<T>(v | (static_cast<T>(std::to_integer<unsigned>(data_[i])) << (CHAR_BIT * i)));
    consume(sizeof(T) * CHAR_BIT);
    return v;
  }
  constexpr auto pop_8() -> std::uint8_t { return pop<std::uint8_t>(); }
  constexpr auto pop_16() -> std::uint16_t { return pop<std::uint16_t>(); }

  /// pointer to the current byte (only meaningful at a byte boundary)
  constexpr auto byte_data() const -> const std::byte*
  {
    assert(bit_offset_ == 0);
    return data_;
  }
};

}  // namespace starflate::huffman
