// amp_deflate.cuh -- BGZF deflate on the device (sm_100a): the step behind the hot path.
//
// The reference writes its trimmed reads through pysam / htslib (AmpliPy.py:911 out_aln.write -> bgzf_write -> zlib deflate at
// level 6 on one host thread); here the rebuilt record stream is compressed in HBM, one warp per BGZF block (<= 0xff00 bytes):
//   deflate_block   LZ77 with a 4-byte hash (most recent occurrence per hash -- atomicMax, so the result does not depend on which lane
//                   wins a collision -- 8 KB of shared memory per warp) plus the distance-1
//                   candidate (runs); the lanes test 32 consecutive positions at a time, each extending its own match word-wise;
//                   the greedy selection then walks the group (warp-uniform) and lane 0 emits the tokens with the FIXED Huffman
//                   code of RFC 1951 3.2.6 (no code construction; tables of ready-made, bit-reversed codes in shared memory)
//   crc32_block     the CRC-32 of the BGZF footer: the lanes take 2 KB segments, the segment CRCs are combined with the
//                   "advance by 2 KB of zeros" operator, whose 32 columns the lanes compute once per warp
// The output is any valid deflate stream (RFC 1951 leaves the choice of matches to the compressor): larger than zlib level 6,
// about the size of level 1, at a small fraction of its time.  A block that does not shrink is stored (BTYPE 00) by the caller.
// Written with the warp primitives of amp_warp.cuh so that tests/emu runs the same source on the CPU (checked against zlib).
#pragma once
#include "amp_bgzf.cuh"

namespace amp {

#define AMPD_HBITS 11
#define AMPD_SEG 2048                 // bytes per lane in the CRC pass (32 lanes cover a whole block)
#define AMPD_MAXBLOCK 0xff00          // BGZF payload limit used by htslib
#define AMPD_SLOT (AMPD_MAXBLOCK + 256)   // bytes of scratch per block for the compressed stream (multiple of 4)

struct DeflateMem { int htab[1 << AMPD_HBITS]; };               // per warp: most recent position of a hash within the block
struct DeflateTables {                                          // per CTA
    uint32_t len_code[260];           // [len 3..258]: fixed code of the length symbol + its extra bits (LSB first) | bit count << 16
    uint32_t crc_tab[256];
    uint16_t lit_code[256];           // bit-reversed fixed code of a literal (8 bits below 144, 9 from there)
};

AMP_HD uint32_t bit_reverse(uint32_t v, int n) { uint32_t r = 0; for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i); return r; }

// built by all threads of the CTA (followed by a block barrier at the caller)
AMP_WD void deflate_tables_init(DeflateTables& T, int tid, int nthreads) {
    for (int b = tid; b < 256; b += nthreads) {
        T.lit_code[b] = (uint16_t)(b < 144 ? bit_reverse(0x30u + (uint32_t)b, 8) : bit_reverse(0x190u + (uint32_t)(b - 144), 9));
        uint32_t c = (uint32_t)b;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        T.crc_tab[b] = c;
    }
    for (int len = tid; len < 260; len += nthreads) {
        uint32_t v = 0;
        if (len >= 3 && len <= 258) {
            int idx = 28;
            while (kLenBase[idx] > len) --idx;
            const int sym = 257 + idx, eb = kLenExtra[idx];
            const int nb = sym <= 279 ? 7 : 8;
            const uint32_t code = sym <= 279 ? bit_reverse((uint32_t)(sym - 256), 7) : bit_reverse(0xC0u + (uint32_t)(sym - 280), 8);
            v = (code | ((uint32_t)(len - kLenBase[idx]) << nb)) | ((uint32_t)(nb + eb) << 16);
        }
        T.len_code[len] = v;
    }
}

// four bytes at any address (two aligned loads; reads up to 7 bytes past p)
AMP_WD uint32_t ld_u32_any(const uint8_t* p) {
    const unsigned long long a = (unsigned long long)p;
    const uint32_t* q = (const uint32_t*)(a & ~3ULL);
    return funnel_r(q[0], q[1], (unsigned)(a & 3ULL) << 3);
}

struct BitWriter { unsigned long long acc; int n; uint32_t* out; int w; };
AMP_WD void bw_put(BitWriter& b, uint32_t code, int len) {      // len <= 31
    b.acc |= (unsigned long long)code << b.n; b.n += len;
    if (b.n >= 32) { b.out[b.w++] = (uint32_t)b.acc; b.acc >>= 32; b.n -= 32; }
}

// length of the match between in[c ...] and in[p ...] (c < p, the first four bytes are known to agree), at most min(258, n - p)
AMP_WD int match_length(const uint8_t* in, int c, int p, int n) {
    const int maxl = n - p < 258 ? n - p : 258;
    int k = 4;
    while (k < maxl) {
        const uint32_t x = ld_u32_any(in + c + k) ^ ld_u32_any(in + p + k);
        if (x) { k += ctz32(x) >> 3; break; }
        k += 4;
    }
    return k < maxl ? k : maxl;
}

// One block by one warp: in[0, n) (n <= AMPD_MAXBLOCK, readable up to in + n + 8) -> a complete deflate stream (one final block,
// fixed Huffman) in out32[0 ...]; returns its length in bytes, or -1 when it would not fit cap_words 32-bit words.
AMP_WD int deflate_block(const uint8_t* in, int n, DeflateMem& M, const DeflateTables& T, uint32_t* out32, int cap_words, int lane) {
    for (int i = lane; i < (1 << AMPD_HBITS); i += 32) M.htab[i] = 0;
    w_sync();
    BitWriter bw; bw.acc = 0; bw.n = 0; bw.out = out32; bw.w = 0;
    if (lane == 0) { bw_put(bw, 1u, 1); bw_put(bw, 1u, 2); }                 // BFINAL = 1, BTYPE = 01
    int cursor = 0;
    for (int base = 0; base < n; base += 32) {
        const int p = base + lane;
        int L = 0, D = 0;
        uint32_t w = 0;
        if (p < n) w = ld_u32_any(in + p);
        const bool hv = p + 4 <= n;
        const uint32_t h = (w * 2654435761u) >> (32 - AMPD_HBITS);
        int c = 0;
        if (hv) c = M.htab[h];
        w_sync();                                                              // every lane has its candidate before the table moves on
        if (hv) atomic_max(&M.htab[h], p);
        if (hv && p >= cursor) {                                               // (positions inside an earlier match only enter the table)
            if (c < p && p - c <= 32768 && ld_u32_any(in + c) == w) { L = match_length(in, c, p, n); D = p - c; }
            if (L < 8 && p >= 1 && ld_u32_any(in + p - 1) == w) {
                const int l1 = match_length(in, p - 1, p, n);
                if (l1 > L) { L = l1; D = 1; }
            }
        }
        // greedy selection over the group, in order; lane 0 writes the tokens
        int s = cursor - base;
        const int lim = n - base < 32 ? n - base : 32;
        while (s < lim) {
            const int Ls = w_shfl(L, s);
            if (Ls >= 4) {
                const int Ds = w_shfl(D, s);
                if (lane == 0) {
                    const uint32_t lc = T.len_code[Ls];
                    bw_put(bw, lc & 0xFFFFu, (int)(lc >> 16));
                    const uint32_t d = (uint32_t)Ds - 1u;
                    uint32_t sym = d, extra = 0; int eb = 0;
                    if (d >= 4u) { const int t = msb32(d); eb = t - 1; sym = 2u * (uint32_t)t + ((d >> eb) & 1u); extra = d & ((1u << eb) - 1u); }
                    bw_put(bw, bit_reverse(sym, 5) | (extra << 5), 5 + eb);
                }
                s += Ls;
            } else {
                const int b = w_shfl((int)(w & 0xFFu), s);
                if (lane == 0) bw_put(bw, T.lit_code[b], b < 144 ? 8 : 9);
                s += 1;
            }
        }
        cursor = base + s;
        if (w_shfl(bw.w, 0) + 40 > cap_words) return -1;                       // a group writes at most 32 tokens of <= 31 bits
    }
    int bytes = 0;
    if (lane == 0) {
        bw_put(bw, 0u, 7);                                                     // end of block (symbol 256)
        bytes = 4 * bw.w + ((bw.n + 7) >> 3);
        if (bw.n > 0) bw.out[bw.w] = (uint32_t)bw.acc;
    }
    return w_shfl(bytes, 0);
}

// column `lane` of the operator that advances the CRC register over AMPD_SEG zero bytes (once per warp)
AMP_WD uint32_t crc_shift_column(const DeflateTables& T, int lane) {
    uint32_t m = 1u << lane;
    for (int i = 0; i < AMPD_SEG; ++i) m = T.crc_tab[m & 0xFFu] ^ (m >> 8);
    return m;
}
AMP_WD uint32_t warp_xor(uint32_t v, int lane) {
    for (int d = 16; d >= 1; d >>= 1) v ^= (uint32_t)w_shfl((int)v, lane ^ d);
    return v;
}
// CRC-32 (zlib's crc32) of in[0, n), n <= 32 * AMPD_SEG, by one warp; mcol = crc_shift_column(T, lane)
AMP_WD uint32_t crc32_block(const uint8_t* in, int n, const DeflateTables& T, uint32_t mcol, int lane) {
    if (n <= 0) return 0u;
    int lo = n - (32 - lane) * AMPD_SEG; const int hi = lo + AMPD_SEG;         // the segments are aligned to the end of the block
    const bool first = lo <= 0 && hi > 0;                                      // the one that holds byte 0 starts from the initial register
    if (lo < 0) lo = 0;
    uint32_t c = first ? 0xFFFFFFFFu : 0u;
    for (int i = lo; i < hi; ++i) c = T.crc_tab[(c ^ in[i]) & 0xFFu] ^ (c >> 8);
    uint32_t acc = 0;
    for (int l = 0; l < 32; ++l) {
        const uint32_t cl = (uint32_t)w_shfl((int)c, l);
        acc = warp_xor(((acc >> lane) & 1u) ? mcol : 0u, lane) ^ cl;
    }
    return acc ^ 0xFFFFFFFFu;
}

}  // namespace amp
