// The reference's own hot-path tests (garymm/starflate src/test/decompress_test.cpp:62-181 and
// the decode-side vectors of huffman/test/{bit_span,decode,table_find_code,
// table_from_symbol_bitsize}_test.cpp), restated against the drop-in headers with a minimal
// expect() in place of boost.ut (un-vendored upstream, absent offline).
//
//   decompress_test host                 host-only cases (no CUDA device needed)
//   decompress_test gpu <golden/bases>   all cases; decompress() runs on cuda:0
#include "huffman/huffman.hpp"
#include "huffman/src/utility.hpp"
#include "src/decompress.hpp"

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

namespace {
int failures = 0;
#define EXPECT(cond)                                                           \
  do {                                                                         \
    if (!(cond)) {                                                             \
      ++failures;                                                              \
      std::fprintf(stderr, "%s:%d: expect failed: %s\n", __FILE__, __LINE__, #cond); \
    }                                                                          \
  } while (0)

auto read_file(const std::string& path) -> std::vector<std::byte>
{
  std::ifstream f{path, std::ios::binary};
  if (!f) {
    std::fprintf(stderr, "cannot open %s\n", path.c_str());
    std::exit(2);
  }
  const std::vector<char> chars((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<std::byte> out(chars.size());
  std::memcpy(out.data(), chars.data(), chars.size());
  return out;
}
auto fnv1a64(std::span<const std::byte> d) -> unsigned long long
{
  unsigned long long h = 1469598103934665603ull;  // (the basis the golden fixtures were hashed with)
  for (auto b : d) h = (h ^ std::to_integer<unsigned>(b)) * 0x100000001b3ull;
  return h;
}

using namespace starflate;
using namespace starflate::huffman::literals;

void test_read_header()  // decompress_test.cpp:62-90
{
  huffman::bit_span empty{nullptr, 0, 0};
  EXPECT(detail::read_header(empty).error() == DecompressStatus::InvalidBlockHeader);

  constexpr auto bad_block_type = huffman::byte_array(0b111);
  huffman::bit_span bad_block_type_span{bad_block_type};
  EXPECT(detail::read_header(bad_block_type_span).error() == DecompressStatus::InvalidBlockHeader);

  constexpr auto fixed = huffman::byte_array(0b010);
  huffman::bit_span fixed_span{fixed};
  auto header = detail::read_header(fixed_span);
  EXPECT(header.has_value());
  EXPECT(!header->final);
  EXPECT(header->type == detail::BlockType::FixedHuffman);
  EXPECT(std::ranges::size(fixed_span) == 5);  // the 3 header bits were consumed

  constexpr auto no_compression = huffman::byte_array(0b001);
  huffman::bit_span no_compression_span{no_compression};
  header = detail::read_header(no_compression_span);
  EXPECT(header.has_value());
  EXPECT(header->final);
  EXPECT(header->type == detail::BlockType::NoCompression);
}

void test_copy_from_before()  // decompress_test.cpp:176-181
{
  auto src_and_dst = huffman::byte_array(1, 2, 0, 0, 0, 0);
  const auto dst_span = std::span<std::byte>{src_and_dst}.subspan(2);
  detail::copy_from_before(2, dst_span.begin(), 3);
  EXPECT(src_and_dst == huffman::byte_array(1, 2, 1, 2, 1, 0));
}

void test_bit_span()  // bit_span_test.cpp: bit order, consume, pop_16 endianness
{
  static constexpr auto data = huffman::byte_array(0b10101010, 0xff, 0x34, 0x12);
  huffman::bit_span bits{data};
  EXPECT(std::ranges::size(bits) == 32);
  const char* expect0 = "01010101";  // LSB first
  for (int i = 0; i < 8; ++i) EXPECT(static_cast<char>(bits[i]) == expect0[i]);
  bits.consume(3);
  EXPECT(static_cast<char>(bits[0]) == '1');
  bits.consume_to_byte_boundary();
  EXPECT(std::ranges::size(bits) == 24);
  EXPECT(bits.pop_8() == 0xff);
  EXPECT(bits.pop_16() == 0x1234);  // little endian
  EXPECT(std::ranges::size(bits) == 0);
  huffman::bit_span off{data.data(), 5, 2};
  EXPECT(static_cast<char>(off[0]) == '0' && static_cast<char>(off[1]) == '1');
}

void test_table_and_decode()
{
  // RFC 1951 §3.2.2 example: lengths (3,3,3,3,3,2,4,4) for A..H
  // (table_from_symbol_bitsize_test.cpp:19-60)
  const huffman::table<char> t{
      huffman::symbol_bitsize,
      std::vector<std::pair<huffman::symbol_span<char>, std::uint8_t>>{
          {huffman::symbol_span<char>{'A', 'E'}, 3}, {huffman::symbol_span<char>{'F'}, 2},
          {huffman::symbol_span<char>{'G', 'H'}, 4}}};
  const std::pair<char, huffman::code> want[] = {{'F', 00_c},  {'A', 010_c},  {'B', 011_c}, {'C', 100_c},
                                                 {'D', 101_c}, {'E', 110_c}, {'G', 1110_c}, {'H', 1111_c}};
  EXPECT(t.size() == 8);
  auto it = t.begin();
  for (const auto& [s, c] : want) {
    EXPECT(it->symbol == s);
    EXPECT(static_cast<huffman::code>(*it) == c);
    ++it;
  }
  // the 288 fixed literal/length codes (:90-149): 0-143 -> 00110000.., 144-255 -> 110010000..,
  // 256-279 -> 0000000.., 280-287 -> 11000000..
  const huffman::table<std::uint16_t, 288> fixed{
      huffman::symbol_bitsize,
      {{{0, 143}, 8}, {{144, 255}, 9}, {{256, 279}, 7}, {{280, 287}, 8}}};
  for (const auto& e : fixed) {
    const std::size_t s = e.symbol;
    const std::size_t v = s < 144 ? 0b00110000 + s : s < 256 ? 0b110010000 + (s - 144)
                          : s < 280 ? (s - 256) : 0b11000000 + (s - 280);
    EXPECT(e.value() == v);
  }
  // find() hints (table_find_code_test.cpp:28-92)
  static constexpr huffman::table table1{huffman::table_contents,
                                         {std::pair{0_c, 'e'}, {10_c, 'i'}, {110_c, 'n'}, {1110_c, 'q'},
                                          {11110_c, '\4'}, {11111_c, 'x'}}};
  static_assert('e' == table1.find(0_c).value()->symbol);
  static_assert('x' == table1.find(11111_c).value()->symbol);
  static_assert(table1.find(1_c).error()->symbol == 'i');
  static_assert(table1.find(111_c).error()->bitsize() == 4);
  static_assert(table1.find(111111_c).error() == table1.end());
  constexpr auto pos1 = table1.find(1_c).error();
  static_assert(table1.find(11_c, pos1).error()->symbol == 'n');
  // decode_one: codes are packed MSB-first into an LSB-first bit stream (decode_test.cpp:155-266)
  static constexpr auto stream = huffman::byte_array(0b00111101);  // bits: 1,0,1,1,1,1,0,0 -> 'i','x'(11111)? no:
  huffman::bit_span bits{stream};
  auto r = huffman::decode_one(table1, bits);  // 1,0 -> 'i'
  EXPECT(r.has_value() && r.symbol() == 'i' && r.encoded_size() == 2);
  bits.consume(r.encoded_size());
  r = huffman::decode_one(table1, bits);       // 1,1,1,1,0 -> '\4'
  EXPECT(r.has_value() && r.symbol() == '\4' && r.encoded_size() == 5);
  bits.consume(r.encoded_size());
  r = huffman::decode_one(table1, bits);       // 0 -> 'e'
  EXPECT(r.has_value() && r.symbol() == 'e' && r.encoded_size() == 1);
  huffman::bit_span none{stream.data(), 0, 0};
  EXPECT(!huffman::decode_one(table1, none).has_value());  // ran out of bits
}

void test_decompress_gpu(const std::string& bases)
{
  {  // "decompress invalid header" (:92-96)
    const auto status = decompress(std::span<const std::byte>{}, std::span<std::byte>{});
    EXPECT(status == DecompressStatus::InvalidBlockHeader);
  }
  {  // "no compression" (:98-134)
    constexpr auto compressed = huffman::byte_array(0b000, 4, 0, ~4, ~0, 'r', 'o', 's', 'e',  //
                                                    0b001, 3, 0, ~3, ~0, 'b', 'u', 'd');
    const std::span<const std::byte> src{compressed};
    constexpr auto expected = huffman::byte_array('r', 'o', 's', 'e', 'b', 'u', 'd');
    std::array<std::byte, expected.size()> dst_array{};
    const std::span<std::byte> dst_too_small{dst_array.data(), dst_array.size() - 1};
    EXPECT(decompress(src, dst_too_small) == DecompressStatus::DstTooSmall);
    const std::span<std::byte> dst{dst_array};
    EXPECT(decompress(src.subspan(0, 5), dst) == DecompressStatus::SrcTooSmall);
    EXPECT(decompress(src, dst) == DecompressStatus::Success);
    EXPECT(std::ranges::equal(dst, expected));
  }
  // "fixed huffman" / "dynamic huffman" (:136-174): starfleet.html, 149618 bytes
  for (const auto& [file, type] : {std::pair{"starfleet_fixed.deflate", detail::BlockType::FixedHuffman},
                                   std::pair{"starfleet_dynamic.deflate", detail::BlockType::DynamicHuffman}}) {
    const std::vector<std::byte> input_bytes = read_file(bases + "/" + file);
    huffman::bit_span input_bits(input_bytes);
    const auto header = detail::read_header(input_bits);
    EXPECT(header.has_value());
    EXPECT(header->type == type);
    std::vector<std::byte> dst(149618);
    const auto status = decompress(input_bytes, dst);  // the contiguous_range overload
    EXPECT(status == DecompressStatus::Success);
    EXPECT(fnv1a64(dst) == 0xe83588b6d41150e9ull);
  }
  {  // batched extension: two streams, one short destination
    constexpr auto a = huffman::byte_array(0x73, 0x04, 0x00);              // fixed: 'A', EOB
    constexpr auto b = huffman::byte_array(0x73, 0x1c, 0x05, 0x00);        // 'A', len 258 dist 1, EOB
    std::array<std::byte, 4> da{};
    std::array<std::byte, 258> db{};
    const std::span<const std::byte> srcs[] = {a, b};
    const std::span<std::byte> dsts[] = {da, db};
    DecompressStatus st[2]{};
    std::size_t wr[2]{};
    EXPECT(decompress_batch(srcs, dsts, st, wr));
    EXPECT(st[0] == DecompressStatus::Success && wr[0] == 1 && da[0] == std::byte{'A'});
    EXPECT(st[1] == DecompressStatus::DstTooSmall && wr[1] == 1);
    // size discovery extension: what the second stream needs, and a failing stream's status
    const auto need = decompressed_size(b);
    EXPECT(need.has_value() && *need == 259);
    constexpr auto bad = huffman::byte_array(0x07);  // BTYPE 3
    const auto none = decompressed_size(bad);
    EXPECT(!none.has_value() && none.error() == DecompressStatus::InvalidBlockHeader);
    const std::vector<std::byte> star = read_file(bases + "/starfleet_dynamic.deflate");
    const auto star_size = decompressed_size(star);
    EXPECT(star_size.has_value() && *star_size == 149618);
  }
  {  // container extension: zlib.compress(b"hello, hello") and the same with a broken Adler-32
    constexpr auto z = huffman::byte_array(0x78, 0x9c, 0xcb, 0x48, 0xcd, 0xc9, 0xc9, 0xd7, 0x51, 0xc8, 0x00, 0x51,
                                           0x00, 0x1c, 0xda, 0x04, 0x75);
    std::array<std::byte, 16> out{};
    std::size_t wr = 0;
    EXPECT(decompress_container(z, out, Container::Zlib, &wr) == ContainerStatus::Success);
    EXPECT(wr == 12 && std::memcmp(out.data(), "hello, hello", 12) == 0);
    EXPECT(decompress_container(z, out, Container::Auto, &wr) == ContainerStatus::Success);
    auto bad = z;
    bad[16] = std::byte{0x5c};
    EXPECT(decompress_container(bad, out, Container::Zlib, &wr) == ContainerStatus::ChecksumMismatch);
    EXPECT(decompress_container(z, out, Container::Gzip, &wr) == ContainerStatus::BadContainer);
  }
  {  // the reference's decompress() is re-entrant ("one stream per core"): eight threads at once,
     // more than the mirror's pool holds contexts, each decoding the test page several times
    const std::vector<std::byte> star = read_file(bases + "/starfleet_dynamic.deflate");
    std::atomic<int> bad{0};
    std::vector<std::thread> threads;
    for (int t = 0; t < 8; ++t)
      threads.emplace_back([&star, &bad, t] {
        for (int k = 0; k < 3; ++k) {
          std::vector<std::byte> out(149618 - (t == 7 ? 5 : 0));
          const auto st = decompress(star, out);
          if (t == 7 ? st != DecompressStatus::DstTooSmall : (st != DecompressStatus::Success || fnv1a64(out) != 0xe83588b6d41150e9ull))
            ++bad;
        }
      });
    for (auto& th : threads) th.join();
    EXPECT(bad.load() == 0);
  }
}
}  // namespace

auto main(int argc, char* argv[]) -> int
{
  const std::string mode = argc > 1 ? argv[1] : "host";
  test_read_header();
  test_copy_from_before();
  test_bit_span();
  test_table_and_decode();
  if (mode == "gpu") test_decompress_gpu(argc > 2 ? argv[2] : "tests/golden/bases");
  std::printf("%s: %d failure(s)\n", mode.c_str(), failures);
  return failures ? 1 : 0;
}
