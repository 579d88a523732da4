// TEST INFRASTRUCTURE ONLY.
//
// Runs the UNMODIFIED reference decompress() over a container of test cases,
// one forked child per case, and reports for each case whether the reference
// returned normally (and with what) or died (sanitizer report / assert abort).
// Built three ways by oracle/Makefile:
//   classify_ndebug  -O2 -DNDEBUG                     status oracle
//   classify_san     -O1 -DNDEBUG -fsanitize=address,undefined   "defined behaviour" classifier
//   classify_assert  -O1 (asserts on)                 what `bazel test` runs upstream
//
// Container format (little endian): u32 n, then n x { u32 src_len, u32 dst_cap, src bytes }.
// Output: one line per case:  <idx> <R|X> <status> <fnv1a64(dst) hex>
//   R = returned normally, X = child died (status/hash printed as 0).
// dst is pre-filled with 0xA5 and hashed over the whole capacity, so bytes the
// reference leaves untouched are part of the fingerprint.
#include "src/decompress.hpp"

#include <sys/wait.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <span>

namespace {
auto fnv1a(const unsigned char* p, std::size_t n) -> std::uint64_t
{
  std::uint64_t h = 1469598103934665603ULL;
  for (std::size_t i = 0; i < n; ++i) {
    h ^= p[i];
    h *= 1099511628211ULL;
  }
  return h;
}
}  // namespace

auto main(int argc, char** argv) -> int
{
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s cases.bin [per-case timeout s]\n", argv[0]);
    return 2;
  }
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  std::uint32_t n = 0;
  if (std::fread(&n, 4, 1, f) != 1) return 2;
  for (std::uint32_t i = 0; i < n; ++i) {
    std::uint32_t hdr[2];
    if (std::fread(hdr, 4, 2, f) != 2) return 2;
    // exact-size heap buffers so that ASan sees any out-of-bounds access
    auto* src = static_cast<unsigned char*>(std::malloc(hdr[0] ? hdr[0] : 1));
    if (hdr[0] && std::fread(src, 1, hdr[0], f) != hdr[0]) return 2;
    unsigned char* exact_src = nullptr;
    if (hdr[0]) {
      exact_src = static_cast<unsigned char*>(std::malloc(hdr[0]));
      std::memcpy(exact_src, src, hdr[0]);
    }
    int fds[2];
    if (pipe(fds) != 0) return 2;
    std::fflush(stdout);
    const pid_t pid = fork();
    if (pid == 0) {
      close(fds[0]);
      // a class-U input can send an unsanitised build into an endless loop: treat as died
      alarm(static_cast<unsigned>(argc > 2 ? std::atoi(argv[2]) : 20));
      // keep sanitizer chatter out of the result stream
      if (!std::freopen("/dev/null", "w", stderr)) _exit(3);
      auto* dst = static_cast<unsigned char*>(std::malloc(hdr[1] ? hdr[1] : 1));
      unsigned char* exact_dst = hdr[1] ? static_cast<unsigned char*>(std::malloc(hdr[1])) : nullptr;
      if (hdr[1]) std::memset(exact_dst, 0xA5, hdr[1]);
      const auto st = starflate::decompress(
          std::span<const std::byte>{reinterpret_cast<const std::byte*>(exact_src), hdr[0]},
          std::span<std::byte>{reinterpret_cast<std::byte*>(exact_dst), hdr[1]});
      std::uint64_t out[2] = {static_cast<std::uint64_t>(st), fnv1a(exact_dst, hdr[1])};
      (void)!write(fds[1], out, sizeof out);
      (void)dst;
      _exit(0);
    }
    close(fds[1]);
    std::uint64_t out[2] = {0, 0};
    const auto got = read(fds[0], out, sizeof out);
    close(fds[0]);
    int wstatus = 0;
    waitpid(pid, &wstatus, 0);
    const bool ok = got == static_cast<ssize_t>(sizeof out) && WIFEXITED(wstatus) &&
                    WEXITSTATUS(wstatus) == 0;
    std::printf("%u %c %llu %016llx\n", i, ok ? 'R' : 'X',
                static_cast<unsigned long long>(ok ? out[0] : 0),
                static_cast<unsigned long long>(ok ? out[1] : 0));
    std::free(src);
    std::free(exact_src);
  }
  std::fclose(f);
  return 0;
}
