// C++23 host side of the drop-in interface: forwards to the CUDA library through the C ABI.
#include "src/decompress.hpp"

#include "starflate_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace starflate {
namespace {

// one context per process, created on first use (the reference has no init/teardown either)
auto context() -> sfb200_ctx*
{
  static std::once_flag once;
  static sfb200_ctx* ctx = nullptr;
  std::call_once(once, [] {
    int device = 0;
    if (const char* e = std::getenv("STARFLATE_B200_DEVICE")) device = std::atoi(e);
    const int rc = sfb200_create(device, &ctx);
    if (rc != SFB200_RC_OK) {
      std::fprintf(stderr,
                   "starflate_b200: cannot create a CUDA context on device %d (rc=%d). "
                   "This library has no CPU decode path.\n",
                   device, rc);
      std::abort();
    }
  });
  return ctx;
}
std::mutex& context_mutex()
{
  static std::mutex m;
  return m;
}

[[noreturn]] void die(const char* what, int rc)
{
  std::fprintf(stderr, "starflate_b200: %s failed (rc=%d): %s\n", what, rc, sfb200_last_error(context()));
  std::abort();
}

}  // namespace

namespace detail {

auto read_header(huffman::bit_span& compressed_bits) -> std::expected<BlockHeader, DecompressStatus>
{
  if (std::ranges::size(compressed_bits) < 3) return std::unexpected{DecompressStatus::InvalidBlockHeader};
  auto it = compressed_bits.begin();
  const bool final_block = static_cast<bool>(*it++);
  unsigned type = static_cast<bool>(*it++) ? 1u : 0u;
  type |= static_cast<bool>(*it++) ? 2u : 0u;
  if (type > 2) return std::unexpected{DecompressStatus::InvalidBlockHeader};
  compressed_bits.consume(3);
  return BlockHeader{final_block, static_cast<BlockType>(type)};
}

void copy_from_before(std::uint16_t distance, std::span<std::byte>::iterator dst, std::uint16_t n)
{
  // byte-serial forward copy: by definition what the LZ77 reference string means (RFC 1951
  // §3.2.3); the device does the same with a period-`distance` gather (csrc/lz_warp.cuh)
  for (std::uint16_t i = 0; i < n; ++i, ++dst) *dst = *(dst - distance);
}

}  // namespace detail

auto decompress(std::span<const std::byte> src, std::span<std::byte> dst) -> DecompressStatus
{
  std::uint8_t status = 0;
  const std::lock_guard lock{context_mutex()};
  const int rc = sfb200_decompress(context(), reinterpret_cast<const std::uint8_t*>(src.data()), src.size(),
                                   reinterpret_cast<std::uint8_t*>(dst.data()), dst.size(), &status, nullptr);
  if (rc != SFB200_RC_OK) die("sfb200_decompress", rc);
  return static_cast<DecompressStatus>(status);
}

auto decompressed_size(std::span<const std::byte> src) -> std::expected<std::size_t, DecompressStatus>
{
  std::uint8_t status = 0;
  std::uint64_t size = 0;
  const std::lock_guard lock{context_mutex()};
  const int rc = sfb200_decompressed_size(context(), reinterpret_cast<const std::uint8_t*>(src.data()), src.size(),
                                          &status, &size);
  if (rc != SFB200_RC_OK) die("sfb200_decompressed_size", rc);
  if (status != 0) return std::unexpected{static_cast<DecompressStatus>(status)};
  return static_cast<std::size_t>(size);
}

auto decompress_container(std::span<const std::byte> src, std::span<std::byte> dst, Container container,
                          std::size_t* written) -> ContainerStatus
{
  std::uint8_t status = 0;
  std::uint64_t wr = 0;
  const std::lock_guard lock{context_mutex()};
  const int rc = sfb200_decompress_container(context(), static_cast<int>(container),
                                             reinterpret_cast<const std::uint8_t*>(src.data()), src.size(),
                                             reinterpret_cast<std::uint8_t*>(dst.data()), dst.size(), &status, &wr);
  if (rc != SFB200_RC_OK) die("sfb200_decompress_container", rc);
  if (written) *written = static_cast<std::size_t>(wr);
  return static_cast<ContainerStatus>(status);
}

auto decompress_batch(std::span<const std::span<const std::byte>> src,
                      std::span<const std::span<std::byte>> dst, std::span<DecompressStatus> status,
                      std::span<std::size_t> written) -> bool
{
  const std::size_t n = src.size();
  if (dst.size() != n || status.size() != n || (!written.empty() && written.size() != n)) return false;
  if (n == 0) return true;
  // pack into the flat layout of the C ABI (16-byte aligned regions)
  std::vector<std::uint64_t> so(n), sl(n), dof(n), dc(n), wr(n);
  std::uint64_t stot = 0, dtot = 0;
  for (std::size_t i = 0; i < n; ++i) {
    so[i] = stot;
    sl[i] = src[i].size();
    stot += (sl[i] + 15) & ~std::uint64_t{15};
    dof[i] = dtot;
    dc[i] = dst[i].size();
    dtot += (dc[i] + 31) & ~std::uint64_t{31};
  }
  std::vector<std::uint8_t> sbuf(stot + 16), dbuf(dtot + 32), st(n);
  for (std::size_t i = 0; i < n; ++i) {
    if (sl[i]) std::memcpy(sbuf.data() + so[i], src[i].data(), sl[i]);
    if (dc[i]) std::memcpy(dbuf.data() + dof[i], dst[i].data(), dc[i]);
  }
  const std::lock_guard lock{context_mutex()};
  const int rc = sfb200_decompress_batch_host(context(), sbuf.data(), sbuf.size(), so.data(), sl.data(),
                                              dbuf.data(), dbuf.size(), dof.data(), dc.data(), st.data(),
                                              wr.data(), n);
  if (rc != SFB200_RC_OK) return false;
  for (std::size_t i = 0; i < n; ++i) {
    if (dc[i]) std::memcpy(dst[i].data(), dbuf.data() + dof[i], dc[i]);
    status[i] = static_cast<DecompressStatus>(st[i]);
    if (!written.empty()) written[i] = static_cast<std::size_t>(wr[i]);
  }
  return true;
}

}  // namespace starflate
