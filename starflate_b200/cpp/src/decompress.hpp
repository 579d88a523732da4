// starflate_b200 — drop-in C++23 interface of the reference's decompress path
// (garymm/starflate src/decompress.hpp:13-71): same namespace, names, enumerator values and
// call shapes, so a caller of the reference compiles against this header unchanged.
// The work is done by the CUDA library behind the C ABI in include/starflate_b200.h
// (nvcc cannot parse C++23, hence the seam).  There is no CPU decode path: without a usable
// CUDA device decompress() terminates the process with a diagnostic.
#pragma once

#include "huffman/huffman.hpp"

#include <cstddef>
#include <cstdint>
#include <expected>
#include <ranges>
#include <span>

namespace starflate {

// numeric values are part of the contract (they cross the C ABI as uint8_t)
enum class DecompressStatus : std::uint8_t
{
  Success,
  Error,
  InvalidBlockHeader,
  NoCompressionLenMismatch,
  DstTooSmall,
  SrcTooSmall,
  InvalidLitOrLen,
  InvalidDistance,
};

namespace detail {

enum class BlockType : std::uint8_t
{
  NoCompression,
  FixedHuffman,
  DynamicHuffman,
};

struct BlockHeader
{
  bool final;
  BlockType type;
};

/// 3-bit block header (BFINAL, BTYPE); consumes it on success (reference :370-385).
auto read_header(huffman::bit_span& compressed_bits) -> std::expected<BlockHeader, DecompressStatus>;

/// Copies n bytes from (dst - distance) to dst, repeating when the ranges overlap
/// (reference :388-398).  @pre dst - distance is valid.
void copy_from_before(std::uint16_t distance, std::span<std::byte>::iterator dst, std::uint16_t n);

}  // namespace detail

/// Decompresses one raw-DEFLATE stream (reference :402-461), on the GPU.
auto decompress(std::span<const std::byte> src, std::span<std::byte> dst) -> DecompressStatus;

template <std::ranges::contiguous_range R>
  requires std::same_as<std::ranges::range_value_t<R>, std::byte>
auto decompress(const R& src, std::span<std::byte> dst)
{
  return decompress(std::span<const std::byte>{src.data(), src.size()}, dst);
}

// ---- extension: size discovery --------------------------------------------------------------
/// The number of bytes decompress(src, dst) produces into a dst that is large enough — which the
/// reference leaves to the caller to know — or the status that ends the stream.  Nothing is
/// stored; one H2D copy of src, one kernel.
auto decompressed_size(std::span<const std::byte> src) -> std::expected<std::size_t, DecompressStatus>;

// ---- extension: zlib (RFC 1950) / gzip (RFC 1952) containers ----------------------------------
enum class Container : std::uint8_t { Raw = 0, Zlib = 1, Gzip = 2, Auto = 3 };
/// DecompressStatus (same values) plus what only a container can get wrong.
enum class ContainerStatus : std::uint8_t
{
  Success = 0,
  Error,
  InvalidBlockHeader,
  NoCompressionLenMismatch,
  DstTooSmall,
  SrcTooSmall,
  InvalidLitOrLen,
  InvalidDistance,
  BadContainer,      // malformed or unsupported header
  ChecksumMismatch,  // Adler-32 / CRC-32 of the output is not the trailer's
  SizeMismatch,      // gzip ISIZE
};
/// decompress() for a whole zlib / gzip container: header, DEFLATE payload, trailer checksum
/// (verified on the device).  `written`, if given, receives the number of bytes produced.
auto decompress_container(std::span<const std::byte> src, std::span<std::byte> dst, Container container,
                          std::size_t* written = nullptr) -> ContainerStatus;

// ---- extension: the batched form the GPU is built for ------------------------------------------
/// Stream i reads src[i] and writes dst[i] (host memory); status[i] / written[i] receive its
/// result.  One call = one H2D copy, two kernels, one D2H copy.  Returns false on an
/// infrastructure (CUDA) failure, which is distinct from any per-stream DecompressStatus.
auto decompress_batch(std::span<const std::span<const std::byte>> src,
                      std::span<const std::span<std::byte>> dst, std::span<DecompressStatus> status,
                      std::span<std::size_t> written) -> bool;

}  // namespace starflate
