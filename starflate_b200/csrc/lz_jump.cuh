// Pass 2 for ONE LARGE stream on sm_100a: LZ77 back-references resolved by POINTER JUMPING over
// the whole output at once (the "two-pass LZ77 resolve" of BASELINE.json's single-stream mode).
//
// lz_warp.cuh gives a stream to one warp, which walks it front to back: right for a batch,
// hopeless for a 1 GB stream.  Here every output byte gets a 32-bit pointer to the byte it is a
// copy of:
//   * a literal (or stored) byte points at itself — it is a ROOT and already holds its value;
//   * byte k of a match (start s, distance d) points at  s - d + (k mod d)  — exactly the byte
//     the reference's copy_from_before (src/decompress.cpp:388-398) would have read, because
//     its forward-overlapping copy is periodic with period d.
// Pointers always point backwards, so following them ends at a root after finitely many hops.
// A round replaces every pointer by the pointer two hops on (ptr[i] <- ptr[ptr[ptr[i]]]), which
// cuts the number of hops to a third: after r rounds every chain of at most 3^r hops is resolved,
// whatever the data (the longest chain a 1 GiB output can hold, 2^30 hops: 19 rounds).  Concurrent updates are benign:
// whatever a thread reads through ptr[p], old or new, is an ancestor of i.
//
// The output can be done in STRIPES, one after the other, each to the end: every target in an
// earlier stripe is then a final byte — a root — which bounds the chains by the stripe, and a small
// stripe's pointers stay in the L2.  Measured on B200 (256 MiB of text): 8 MiB stripes 9.9 ms, one
// stripe 6.0 ms — a round over a small stripe is a short kernel bound by launch and memory latency,
// not by bandwidth — so the default is ONE stripe (SFB200_JUMP_STRIPE_MB sets another size).
//
// A stripe is cut into tiles of JUMP_TILE bytes, one warp per tile and round.  A tile in which a
// round changed nothing has only roots as targets; roots never change, so the tile is FINAL: the
// warp gathers its bytes (dst[i] <- dst[ptr[i]]; only roots are read and no root is rewritten
// with a different value) and marks it done.  Most bytes of text are final after a few rounds
// but the deepest chain of a tile takes ten and more, so a round only touches the pointers that
// were not final after the one before (a bit per 4 bytes).  The host enqueues JUMP_MAX_ROUNDS launches; a
// launch whose predecessor left nothing to do exits at once (no host round trip).
//
// Positions are 32-bit offsets in the 128-byte aligned view of the stream (as in lz_warp.cuh):
// the batch precondition (sizes < 0xffffff00) makes them fit.
#pragma once

#include "lz_warp.cuh"

namespace sfb {

constexpr uint32_t JUMP_TILE = 1024;      // bytes of output per warp and round
constexpr int JUMP_MAX_ROUNDS = 22;       // 3^21 hops > 2^32 positions, + 1 to see that nothing moved
constexpr int JUMP_THREADS = 256;

struct JumpArgs {
  uint8_t* dst_base;  // 128-byte aligned
  uint64_t dst_delta;
  const uint64_t* dst_off;
  const uint64_t* written;
  const uint32_t* match_bits;
  uint64_t idx;         // the stream
  uint32_t* ptr;        // one per byte of the view [0, q + written) rounded up to a tile
  uint32_t* tile_done;  // one per tile, zeroed before the first round
  uint32_t* open;       // JUMP_TILE / 128 words per tile: groups of 4 pointers that are not final yet
  uint32_t* todo;       // this stripe's [JUMP_MAX_ROUNDS + 1], zeroed; todo[r] != 0: round r left tiles open
  uint32_t round;       // 0 = set up the pointers, then 1 .. JUMP_MAX_ROUNDS
  uint32_t tile_lo;     // the stripe: tiles [tile_lo, tile_hi); everything before it is final already
  uint32_t tile_hi;
};

// Round 0: the pointers.  A warp takes a tile, finds the match that reaches into it from the
// left (the last match head in the 258 bytes before the tile, if its length carries it that far)
// and then walks the tile's chunks the way lz_resolve_kernel does.
__global__ void __launch_bounds__(JUMP_THREADS) lz_jump_init_kernel(const JumpArgs a)
{
  constexpr unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int next_lane = static_cast<int>((lane + 1u) & 31u);
  const uint32_t bit_sh = 4u * (lane & 7u);
  const uint64_t off = a.dst_off[a.idx] + a.dst_delta;
  const uint64_t wr = a.written[a.idx];
  if (wr == 0) return;
  uint8_t* const base = a.dst_base + (off & ~127ull);
  const uint32_t q = static_cast<uint32_t>(off & 127u);
  const uint32_t end = q + static_cast<uint32_t>(wr);
  const uint32_t* const pw = reinterpret_cast<const uint32_t*>(base) + lane;
  const uint32_t* const bm = a.match_bits + ((off & ~127ull) >> 5);
  const uint32_t all_tiles = (end + JUMP_TILE - 1u) / JUMP_TILE;
  const uint32_t n_tiles = all_tiles < a.tile_hi ? all_tiles : a.tile_hi;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t tile = a.tile_lo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); tile < n_tiles; tile += warps) {
    const uint32_t T = tile * JUMP_TILE;
    // ---- the match carried in: last head in [T - 288, T), nine bitmap words, lanes 0..8 -----------
    uint32_t c_o = 0, c_end = 0, c_d = 1;
    if (T > q) {
      const uint32_t w_hi = T >> 5;  // first word of the tile
      uint32_t best = 0;             // (bit index + 1) of my word's top head, 0: none
      if (lane < 9u && w_hi > lane) {
        const uint32_t wi = w_hi - 1u - lane;
        uint32_t w = bm[wi];
        if (32u * wi < q) w &= q >= 32u * wi + 32u ? 0u : ~((1u << (q - 32u * wi)) - 1u);
        if (w) best = 32u * wi + (31u - static_cast<uint32_t>(__clz(static_cast<int>(w)))) + 1u;
      }
      best = __reduce_max_sync(FULL, best);
      if (best) {
        const uint32_t h = best - 1u;
        const uint32_t d = static_cast<uint32_t>(base[h]) | (static_cast<uint32_t>(base[h + 1u]) << 8) |
                           (static_cast<uint32_t>(base[h + 2u]) << 16);
        c_o = h;
        c_end = h + (d & 255u) + 3u;
        c_d = (d >> 8) + 1u;
      }
    }
    auto load_w = [&](uint32_t P) -> uint32_t {
      const uint32_t wp = P + 4u * lane;
      return (wp + 4u > q && wp < end) ? pw[P >> 2] : 0u;
    };
    const uint32_t t_stop = T + JUMP_TILE < end ? T + JUMP_TILE : end;
    uint32_t cw = load_w(T);
    for (uint32_t P = T; P < t_stop; P += 128u) {
      const uint32_t ncw = load_w(P + 128u);
      const uint32_t mw = bm[(P >> 5) + (lane >> 3)];
      const uint32_t wp = P + 4u * lane;
      uint32_t vm = 15u;
      if (!(P >= q && P + 128u <= end)) {
        const uint32_t lo = q > wp ? (q - wp < 4u ? q - wp : 4u) : 0u;
        const uint32_t hi = end > wp ? (end - wp < 4u ? end - wp : 4u) : 0u;
        vm = hi > lo ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
      }
      const uint32_t hb4 = (mw >> bit_sh) & vm;
      const uint32_t t = __shfl_sync(FULL, cw, next_lane);
      const uint32_t u = __shfl_sync(FULL, ncw, 0);
      const uint32_t nx = lane == 31u ? u : t;
      const uint32_t hl = 31u - static_cast<uint32_t>(__clz(static_cast<int>(hb4 | 1u)));
      const uint32_t own_pack = (4u * lane + hl) | ((lz_funnel(cw, nx, 8u * hl) & 0xffffffu) << 7);
      const uint32_t hm = __ballot_sync(FULL, hb4 != 0);
      const uint32_t below = hm & lt_mask;
      const uint32_t sl = 31u - static_cast<uint32_t>(__clz(static_cast<int>(below | 1u)));
      const uint32_t in_pack = __shfl_sync(FULL, own_pack, static_cast<int>(sl));
      uint32_t t_o = c_o, t_end = c_end, t_d = c_d;
      if (below) {
        t_o = P + (in_pack & 127u);
        t_end = t_o + ((in_pack >> 7) & 255u) + 3u;
        t_d = (in_pack >> 15) + 1u;
      }
      if (hm) {
        const uint32_t top = 31u - static_cast<uint32_t>(__clz(static_cast<int>(hm)));
        const uint32_t pk = __shfl_sync(FULL, own_pack, static_cast<int>(top));
        c_o = P + (pk & 127u);
        c_end = c_o + ((pk >> 7) & 255u) + 3u;
        c_d = (pk >> 15) + 1u;
      }
      uint32_t src[4];
#pragma unroll
      for (uint32_t b = 0; b < 4; ++b) {
        const uint32_t p = wp + b;
        if ((hb4 >> b) & 1u) {
          const uint32_t d = lz_funnel(cw, nx, 8u * b) & 0xffffffu;
          t_o = p;
          t_end = p + (d & 255u) + 3u;
          t_d = (d >> 8) + 1u;
        }
        const bool cov = ((vm >> b) & 1u) && p < t_end;
        uint32_t k = p - t_o;
        if (cov && k >= t_d) k = lz_mod_small(k, t_d);  // (k < 258)
        src[b] = cov ? t_o - t_d + k : p;
      }
      *reinterpret_cast<uint4*>(a.ptr + wp) = make_uint4(src[0], src[1], src[2], src[3]);
      cw = ncw;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.tile_lo < n_tiles) a.todo[0] = 1u;
}

// Rounds 1 .. JUMP_MAX_ROUNDS.  Per tile, 8 words of `open` bits say which 16-byte groups of
// pointers (4 output bytes) still had a non-root target after the previous round; only those are
// visited, and each takes two hops at once (ptr <- ptr[ptr[ptr]]).
__global__ void __launch_bounds__(JUMP_THREADS) lz_jump_round_kernel(const JumpArgs a)
{
  constexpr unsigned FULL = 0xffffffffu;
  constexpr uint32_t ROWS = JUMP_TILE / 128u;
  if (a.todo[a.round - 1u] == 0u) return;  // everything was final before this round
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t off = a.dst_off[a.idx] + a.dst_delta;
  const uint64_t wr = a.written[a.idx];
  if (wr == 0) return;
  uint8_t* const base = a.dst_base + (off & ~127ull);
  const uint32_t q = static_cast<uint32_t>(off & 127u);
  const uint32_t end = q + static_cast<uint32_t>(wr);
  const uint32_t all_tiles = (end + JUMP_TILE - 1u) / JUMP_TILE;
  const uint32_t n_tiles = all_tiles < a.tile_hi ? all_tiles : a.tile_hi;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const bool first = a.round == 1u;
  const uint32_t lo = a.tile_lo * JUMP_TILE;  // targets below this are final bytes of earlier stripes: roots
  bool open_any = false;
  for (uint32_t tile = a.tile_lo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); tile < n_tiles; tile += warps) {
    if (a.tile_done[tile]) continue;  // (warp-uniform)
    const uint32_t T = tile * JUMP_TILE;
    uint32_t* const om = a.open + static_cast<size_t>(tile) * ROWS;
    // my row's mask of open groups (lane j < ROWS holds row j's)
    const uint32_t mine = first ? FULL : (lane < ROWS ? om[lane] : 0u);
    auto hop = [&](uint32_t t) { return t < lo ? t : a.ptr[t]; };
    // All rows' loads are issued before anything is stored (the compiler must assume that a store
    // to ptr[] changes what the next row reads, and would otherwise run the rows one after the
    // other: eight chains of three dependent loads instead of one).
    uint4 p[ROWS], g[ROWS];
    bool act[ROWS], ch[ROWS];
#pragma unroll
    for (uint32_t j = 0; j < ROWS; ++j) {
      const uint32_t mask = __shfl_sync(FULL, mine, static_cast<int>(j));
      const uint32_t wp = T + 128u * j + 4u * lane;
      act[j] = ((mask >> lane) & 1u) && wp < end;
      p[j] = make_uint4(wp, wp + 1u, wp + 2u, wp + 3u);
      if (act[j]) p[j] = *reinterpret_cast<const uint4*>(a.ptr + wp);
    }
#pragma unroll
    for (uint32_t j = 0; j < ROWS; ++j) {
      g[j] = p[j];
      if (act[j]) {
        g[j].x = hop(p[j].x);
        g[j].y = hop(p[j].y);
        g[j].z = hop(p[j].z);
        g[j].w = hop(p[j].w);
      }
    }
#pragma unroll
    for (uint32_t j = 0; j < ROWS; ++j) {
      ch[j] = (g[j].x != p[j].x) | (g[j].y != p[j].y) | (g[j].z != p[j].z) | (g[j].w != p[j].w);
      if (ch[j]) {
        g[j].x = hop(g[j].x);
        g[j].y = hop(g[j].y);
        g[j].z = hop(g[j].z);
        g[j].w = hop(g[j].w);
      }
    }
    uint32_t still = 0;  // row masks after this round, gathered in lane j
#pragma unroll
    for (uint32_t j = 0; j < ROWS; ++j) {
      const uint32_t wp = T + 128u * j + 4u * lane;
      if (ch[j]) *reinterpret_cast<uint4*>(a.ptr + wp) = g[j];
      const uint32_t nm = __ballot_sync(FULL, ch[j]);
      if (lane == j) still = nm;
    }
    const bool open_tile = __any_sync(FULL, still != 0u);
    if (open_tile) {
      if (lane < ROWS) om[lane] = still;
      open_any = true;
      continue;
    }
    // final: every target is a root.  (p[] holds the final pointers of the groups that were still
    // open; the others are read again.)
    uint32_t v[ROWS];
#pragma unroll
    for (uint32_t j = 0; j < ROWS; ++j) {
      const uint32_t wp = T + 128u * j + 4u * lane;
      if (!act[j] && wp < end) p[j] = *reinterpret_cast<const uint4*>(a.ptr + wp);
    }
#pragma unroll
    for (uint32_t j = 0; j < ROWS; ++j) {
      const uint32_t wp = T + 128u * j + 4u * lane;
      v[j] = 0;
      if (wp < end && wp + 4u > q)
        v[j] = static_cast<uint32_t>(base[p[j].x]) | (static_cast<uint32_t>(base[p[j].y]) << 8) |
               (static_cast<uint32_t>(base[p[j].z]) << 16) | (static_cast<uint32_t>(base[p[j].w]) << 24);
    }
#pragma unroll
    for (uint32_t j = 0; j < ROWS; ++j) {
      const uint32_t wp = T + 128u * j + 4u * lane;
      if (wp >= end || wp + 4u <= q) continue;
      if (wp >= q && wp + 4u <= end) {
        *reinterpret_cast<uint32_t*>(base + wp) = v[j];
      } else {  // first / last word of the stream: only our bytes
#pragma unroll
        for (uint32_t b = 0; b < 4; ++b)
          if (wp + b >= q && wp + b < end) base[wp + b] = static_cast<uint8_t>(v[j] >> (8u * b));
      }
    }
    if (lane == 0) a.tile_done[tile] = 1u;
  }
  if (__any_sync(FULL, open_any) && lane == 0) a.todo[a.round] = 1u;
}

}  // namespace sfb
