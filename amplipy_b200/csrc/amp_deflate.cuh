// amp_deflate.cuh -- BGZF deflate on the device (sm_100a): the step behind the hot path.
//
// The reference writes its trimmed reads through pysam / htslib (AmpliPy.py:911 out_aln.write -> bgzf_write -> zlib deflate at
// level 6 on one host thread); here the rebuilt record stream is compressed in HBM, one warp per BGZF block (<= 0xff00 bytes):
//   deflate_block   LZ77 with a 4-byte hash (most recent occurrence per hash -- atomicMax, so the result does not depend on which lane
//                   wins a collision -- 8 KB of shared memory per warp) plus the distance-1
//                   candidate (runs); the lanes test 32 consecutive positions at a time, each extending its own match word-wise;
//                   the greedy selection then walks the group (warp-uniform) and lane 0 records the tokens; every AMPD_TOKCAP tokens
//                   become one deflate block with a Huffman code of their own (huff_lengths: frequency ranking across the lanes, then
//                   the in-place minimum-redundancy algorithm, length-limited by halving) or the fixed code of RFC 1951 3.2.6,
//                   whichever is shorter
//   crc32_block     the CRC-32 of the BGZF footer: the lanes take 2 KB segments, the segment CRCs are combined with the
//                   "advance by 2 KB of zeros" operator, whose 32 columns the lanes compute once per warp
// The output is any valid deflate stream (RFC 1951 leaves the choice of matches to the compressor): larger than zlib level 6,
// a little smaller than level 1, at a small fraction of its time.  A block that does not shrink is stored (BTYPE 00) by the caller.
// Written with the warp primitives of amp_warp.cuh so that tests/emu runs the same source on the CPU (checked against zlib).
#pragma once
#include "amp_bgzf.cuh"

namespace amp {

#define AMPD_HBITS 11
#define AMPD_SEG 2048                 // bytes per lane in the CRC pass (32 lanes cover a whole block)
#define AMPD_MAXBLOCK 0xff00          // BGZF payload limit used by htslib
#define AMPD_SLOT (AMPD_MAXBLOCK + 256)   // bytes of scratch per block for the compressed stream (multiple of 4)

#define AMPD_TOKCAP 8192              // tokens of one deflate block (a BGZF block becomes a few of them, each with its own code)
#define AMPD_SCRATCH (AMPD_TOKCAP + 512)   // words of global scratch per warp: the tokens, then the code construction's work arrays
struct DeflateMem {                                             // per warp (9.2 KB: 24 warps per SM)
    int htab[1 << AMPD_HBITS];                                  // most recent position of a hash within the BGZF block
    uint16_t lfreq[288], dfreq[32];                             // symbol frequencies of the open deflate block; its codes while it is written
    uint8_t llen[288], dlen[32], clen[20];                      // code lengths: literal / length, distance, code-length code
    uint16_t cfreq[20];                                         // frequencies of the code lengths themselves; then the code-length code
};
struct HuffWork { uint32_t* a; uint16_t* order; };              // in the warp's global scratch: 288 words (frequencies -> lengths in place), 288 symbols by ascending frequency
// token: literal = the byte; match = 1 << 31 | length symbol index << 26 | distance symbol << 21 | length extra << 16 | distance extra
struct DeflateTables {                                          // per CTA
    uint32_t len_code[260];           // [len 3..258]: fixed code of the length symbol + its extra bits (LSB first) | bit count << 16
    uint16_t len_info[260];           // [len 3..258]: index of the length symbol (0..28) | its extra-bit value << 8
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
            T.len_info[len] = (uint16_t)(idx | ((len - kLenBase[idx]) << 8));
        } else T.len_info[len] = 0;
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

// ---- code construction ---------------------------------------------------------------------------------------------------------
// Code lengths (at most maxlen) of the symbols with freq[i] > 0 among n_sym, into len[] (0 = unused); at least two symbols get a
// code (as zlib does, so that no special cases are left for the decoder).  All lanes call; freq / len in shared memory, W in the
// warp's global scratch (touched a few times per 8 k tokens).
// Ranking by (frequency, symbol) across the lanes, then the in-place minimum-redundancy algorithm of Moffat and Katajainen on the
// sorted frequencies by lane 0; a code that comes out too long is recomputed from halved frequencies.
AMP_WD void huff_lengths(uint16_t* freq, int n_sym, int maxlen, uint8_t* len, const HuffWork& W, int lane) {
    if (lane == 0) {
        int used = 0, which = -1;
        for (int i = 0; i < n_sym; ++i) if (freq[i]) { ++used; which = i; }
        if (used == 0) { freq[0] = 1; freq[1] = 1; }
        else if (used == 1) freq[which == 0 ? 1 : 0] = 1;
    }
    w_sync();
    int mine = 0;
    for (int i = lane; i < n_sym; i += 32) {
        len[i] = 0;
        const unsigned f = freq[i];
        if (!f) continue;
        int rank = 0;
        for (int j = 0; j < n_sym; ++j) { const unsigned g = freq[j]; if (g && (g < f || (g == f && j < i))) ++rank; }
        W.order[rank] = (uint16_t)i;
        ++mine;
    }
    const int n = w_add(mine);
    w_sync();
    if (lane == 0) {
        uint32_t* A = W.a;
        for (int shift = 0;; ++shift) {
            for (int k = 0; k < n; ++k) { const uint32_t f = (uint32_t)freq[W.order[k]] >> shift; A[k] = f ? f : 1u; }
            // first pass, left to right: parent pointers
            A[0] += A[1];
            int root = 0, leaf = 2;
            for (int next = 1; next < n - 1; ++next) {
                if (leaf >= n || A[root] < A[leaf]) { A[next] = A[root]; A[root++] = (uint32_t)next; } else A[next] = A[leaf++];
                if (leaf >= n || (root < next && A[root] < A[leaf])) { A[next] += A[root]; A[root++] = (uint32_t)next; } else A[next] += A[leaf++];
            }
            // second pass, right to left: depths of the internal nodes
            A[n - 2] = 0;
            for (int next = n - 3; next >= 0; --next) A[next] = A[A[next]] + 1;
            // third pass, right to left: depths of the leaves
            int avbl = 1, used = 0, dpth = 0, next = n - 1;
            root = n - 2;
            while (avbl > 0) {
                while (root >= 0 && (int)A[root] == dpth) { ++used; --root; }
                while (avbl > used) { A[next--] = (uint32_t)dpth; --avbl; }
                avbl = 2 * used; ++dpth; used = 0;
            }
            if ((int)A[0] <= maxlen) break;
        }
        for (int k = 0; k < n; ++k) len[W.order[k]] = (uint8_t)A[k];
    }
    w_sync();
}
// canonical codes of the lengths len[0, n), bit-reversed (deflate packs codes starting from their most significant bit); lane 0
AMP_WD void huff_codes(const uint8_t* len, int n, uint16_t* code) {
    int count[16], next[16];
    for (int l = 0; l < 16; ++l) count[l] = 0;
    for (int i = 0; i < n; ++i) ++count[len[i] & 15];
    count[0] = 0;
    int c = 0;
    for (int l = 1; l < 16; ++l) { c = (c + count[l - 1]) << 1; next[l] = c; }
    for (int i = 0; i < n; ++i) code[i] = len[i] ? (uint16_t)bit_reverse((uint32_t)next[len[i]]++, len[i]) : (uint16_t)0;
}

// The open deflate block (ntok tokens in tok[], their frequencies in M) written out: the code is built from the frequencies, or the
// fixed code of RFC 1951 3.2.6 is used where that is shorter; `last` sets BFINAL.  xbits = extra bits of the tokens.  All lanes call;
// returns false (warp-uniform) when the block does not fit cap_words.  Leaves the frequencies cleared.
AMP_WD bool deflate_flush(DeflateMem& M, const HuffWork& W, const uint32_t* tok, int ntok, int xbits, bool last, BitWriter& bw, int cap_words, int lane) {
    if (lane == 0) ++M.lfreq[256];                                             // end of block
    w_sync();
    huff_lengths(M.lfreq, 286, 15, M.llen, W, lane);
    huff_lengths(M.dfreq, 30, 15, M.dlen, W, lane);
    int hlit = 0, hdist = 0;
    bool dyn = false, fits = true;
    if (lane == 0) {
        hlit = 286; while (hlit > 257 && !M.llen[hlit - 1]) --hlit;
        hdist = 30; while (hdist > 1 && !M.dlen[hdist - 1]) --hdist;
        for (int v = 0; v < 19; ++v) M.cfreq[v] = 0;
        for (int i = 0; i < hlit; ++i) ++M.cfreq[M.llen[i]];
        for (int i = 0; i < hdist; ++i) ++M.cfreq[M.dlen[i]];
    }
    w_sync();
    huff_lengths(M.cfreq, 19, 7, M.clen, W, lane);
    if (lane == 0) {
        long long cdyn = 3 + 14 + 19 * 3, cfix = 3;
        for (int v = 0; v < 19; ++v) cdyn += (long long)M.cfreq[v] * M.clen[v];
        for (int i = 0; i < 286; ++i) { cdyn += (long long)M.lfreq[i] * M.llen[i]; cfix += (long long)M.lfreq[i] * (i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8); }
        for (int i = 0; i < 30; ++i) { cdyn += (long long)M.dfreq[i] * M.dlen[i]; cfix += (long long)M.dfreq[i] * 5; }
        dyn = cdyn < cfix;
        const long long bits = (dyn ? cdyn : cfix) + xbits;
        fits = 4LL * bw.w + (bits >> 3) + 64 <= 4LL * cap_words;
        if (fits) {
            uint16_t* lcode = M.lfreq; uint16_t* dcode = M.dfreq; uint16_t* ccode = M.cfreq;     // the frequencies have served: their arrays hold the codes now
            bw_put(bw, last ? 1u : 0u, 1);
            if (dyn) {
                M.llen[286] = 0; M.llen[287] = 0; M.dlen[30] = 0; M.dlen[31] = 0;     // (symbols that cannot occur; huff_lengths leaves them alone)
                bw_put(bw, 2u, 2);
                bw_put(bw, (uint32_t)(hlit - 257), 5); bw_put(bw, (uint32_t)(hdist - 1), 5); bw_put(bw, 15u, 4);
                for (int k = 0; k < 19; ++k) bw_put(bw, M.clen[kClOrder[k]], 3);
                huff_codes(M.clen, 19, ccode);
                for (int i = 0; i < hlit; ++i) bw_put(bw, ccode[M.llen[i]], M.clen[M.llen[i]]);
                for (int i = 0; i < hdist; ++i) bw_put(bw, ccode[M.dlen[i]], M.clen[M.dlen[i]]);
            } else {
                bw_put(bw, 1u, 2);
                for (int i = 0; i < 288; ++i) M.llen[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                for (int i = 0; i < 32; ++i) M.dlen[i] = 5;
            }
            huff_codes(M.llen, 288, lcode);
            huff_codes(M.dlen, 32, dcode);
            for (int t = 0; t < ntok; ++t) {
                const uint32_t k = tok[t];
                if (!(k >> 31)) { bw_put(bw, lcode[k], M.llen[k]); continue; }
                const uint32_t idx = (k >> 26) & 31u, ds = (k >> 21) & 31u, lx = (k >> 16) & 31u, dx = k & 0x1FFFu;
                const int ll = M.llen[257 + idx], dl = M.dlen[ds];
                bw_put(bw, lcode[257 + idx] | (lx << ll), ll + kLenExtra[idx]);
                bw_put(bw, dcode[ds] | (dx << dl), dl + (ds < 4u ? 0 : (int)(ds >> 1) - 1));
            }
            bw_put(bw, lcode[256], M.llen[256]);
        }
    }
    w_sync();
    for (int i = lane; i < 288; i += 32) M.lfreq[i] = 0;
    if (lane < 32) M.dfreq[lane] = 0;
    w_sync();
    return w_shfl(fits ? 1 : 0, 0) != 0;
}

// One BGZF block by one warp: in[0, n) (n <= AMPD_MAXBLOCK, readable up to in + n + 8) -> a complete deflate stream in out32[0 ...]
// (deflate blocks of up to AMPD_TOKCAP tokens, each with the shorter of its own and the fixed code); tok = AMPD_SCRATCH words of
// scratch in global memory.  Returns the stream's length in bytes, or -1 when it would not fit cap_words 32-bit words.
AMP_WD int deflate_block(const uint8_t* in, int n, DeflateMem& M, const DeflateTables& T, uint32_t* out32, int cap_words, uint32_t* tok, int lane) {
    HuffWork W; W.a = tok + AMPD_TOKCAP; W.order = (uint16_t*)(tok + AMPD_TOKCAP + 288);
    for (int i = lane; i < (1 << AMPD_HBITS); i += 32) M.htab[i] = 0;
    for (int i = lane; i < 288; i += 32) M.lfreq[i] = 0;
    M.dfreq[lane] = 0;
    w_sync();
    BitWriter bw; bw.acc = 0; bw.n = 0; bw.out = out32; bw.w = 0;
    int cursor = 0, ntok = 0, xbits = 0;
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
        // greedy selection over the group, in order; lane 0 records the tokens
        int s = cursor - base;
        const int lim = n - base < 32 ? n - base : 32;
        while (s < lim) {
            const int Ls = w_shfl(L, s);
            if (Ls >= 4) {
                const int Ds = w_shfl(D, s);
                if (lane == 0) {
                    const uint32_t info = T.len_info[Ls], idx = info & 0xFFu, lx = info >> 8;
                    const uint32_t d = (uint32_t)Ds - 1u;
                    uint32_t sym = d, extra = 0; int eb = 0;
                    if (d >= 4u) { const int t = msb32(d); eb = t - 1; sym = 2u * (uint32_t)t + ((d >> eb) & 1u); extra = d & ((1u << eb) - 1u); }
                    tok[ntok++] = 0x80000000u | (idx << 26) | (sym << 21) | (lx << 16) | extra;
                    ++M.lfreq[257 + idx]; ++M.dfreq[sym];
                    xbits += kLenExtra[idx] + eb;
                }
                s += Ls;
            } else {
                const int b = w_shfl((int)(w & 0xFFu), s);
                if (lane == 0) { tok[ntok++] = (uint32_t)b; ++M.lfreq[b]; }
                s += 1;
            }
        }
        cursor = base + s;
        if (w_shfl(ntok, 0) > AMPD_TOKCAP - 32) {                              // (a group adds at most 32 tokens)
            if (!deflate_flush(M, W, tok, ntok, xbits, false, bw, cap_words, lane)) return -1;
            ntok = 0; xbits = 0;
        }
    }
    if (!deflate_flush(M, W, tok, ntok, xbits, true, bw, cap_words, lane)) return -1;
    int bytes = 0;
    if (lane == 0) {
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
