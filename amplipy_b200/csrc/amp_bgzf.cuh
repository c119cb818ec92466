// amp_bgzf.cuh -- BGZF / BAM decode on the device (sm_100a): the step in front of the hot path.
//
// The reference gets its records from pysam / htslib (AmpliPy.py:296-360 create_AlignmentFile_objects, 896 iteration); here the
// compressed file crosses PCIe as it is (about a fifth of the decoded struct-of-arrays bytes) and is decoded in HBM:
//   inflate_block   one warp per BGZF block (RFC 1951): lane 0 reads bits and decodes Huffman symbols -- literals are stored as
//                   they come -- and hands every match (length, distance) to the warp, whose lanes copy it together; the code
//                   tables of a block (a 10-bit direct table + canonical lists for longer codes) live in the warp's shared memory
//   bam_count       one thread per BGZF block walks its record chain: htslib never lets a record straddle a block boundary
//                   (bgzf_flush_try in bam_write1), so every block starts on a record; a block that does not end on one raises an
//                   error and the host falls back to its own decoder
//   bam_scatter     one warp per block: records -> pos / flag / tlen / packed CIGAR / 4-bit seq / qual arrays at the offsets the
//                   prefix sums over the per-block totals give
// Written with the warp primitives of amp_warp.cuh so that tests/emu runs the same source on the CPU (against zlib).
#pragma once
#include "amp_warp.cuh"

namespace amp {

#define AMPZ_LBITS 10                 // direct table of the literal / length code
#define AMPZ_DBITS 8                  // direct table of the distance code
#define AMPZ_E_DATA 1                 // malformed deflate stream
#define AMPZ_E_SIZE 2                 // output does not match ISIZE
#define AMPZ_E_ALIGN 4                // a BAM record straddles a block boundary (or a record header does not fit its record)

// RFC 1951 3.2.5: base values / extra bits of the length and distance symbols, order of the code-length code lengths
#if defined(__CUDACC__)
#define AMPZ_TABLE __device__ __constant__
#else
#define AMPZ_TABLE static const
#endif
AMPZ_TABLE uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
AMPZ_TABLE uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
AMPZ_TABLE uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
AMPZ_TABLE uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
AMPZ_TABLE uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct HuffLists { uint16_t count[16]; uint16_t symbol[288]; };          // canonical code: symbols ordered by (length, symbol)
#define AMPZ_RING 64                  // words of the compressed stream kept in shared memory ahead of the bit reader
#define AMPZ_WIN 2048                 // bytes of output kept in shared memory (circular): source of the near matches
struct InflateMem {                                                       // per warp, in shared memory
    uint16_t ltab[1 << AMPZ_LBITS];                                       // (symbol << 4) | length, 0 = longer than the table
    uint16_t dtab[1 << AMPZ_DBITS];
    HuffLists lit, dist;
    uint8_t lens[320];
    uint32_t ring[AMPZ_RING];
    uint8_t win[AMPZ_WIN];
};

// Bit reader of lane 0.  Words come from the ring in shared memory when they are there (the warp keeps it ahead of the reader at
// every point where its lanes meet) and from global memory otherwise, so correctness never depends on the ring.
struct BitReader {
    const uint32_t* words;            // the stream's aligned words in global memory
    const uint32_t* ring;
    int nwords;                       // words that hold stream bytes
    int widx;                         // next word to take
    int ring_hi;                      // the ring holds words [ring_hi - AMPZ_RING, ring_hi)
    unsigned long long buf; int cnt;  // cnt valid bits in buf
};
AMP_WD uint32_t br_word(const BitReader& b, int idx) {
    if (idx >= b.nwords) return 0u;                                       // zeros past the end of the input
    if (idx < b.ring_hi && idx >= b.ring_hi - AMPZ_RING) return b.ring[idx & (AMPZ_RING - 1)];
    return b.words[idx];
}
AMP_WD void br_init(BitReader& b, const uint8_t* p, const uint8_t* end, const uint32_t* ring) {
    const uintptr_t a = (uintptr_t)p;
    b.words = (const uint32_t*)(a & ~(uintptr_t)3);
    b.ring = ring; b.ring_hi = 0;
    b.nwords = (int)((((uintptr_t)end + 3) & ~(uintptr_t)3) - (a & ~(uintptr_t)3)) >> 2;
    const int skip = (int)(a & 3) * 8;
    b.buf = (unsigned long long)(b.nwords > 0 ? b.words[0] : 0u) >> skip; b.cnt = 32 - skip; b.widx = 1;
}
AMP_WD void br_refill(BitReader& b) {               // at least 32 valid bits afterwards
    if (b.cnt <= 32) {
        b.buf |= (unsigned long long)br_word(b, b.widx) << b.cnt;
        ++b.widx; b.cnt += 32;
    }
}
AMP_WD unsigned br_peek(const BitReader& b, int n) { return (unsigned)(b.buf & ((1ULL << n) - 1ULL)); }
AMP_WD void br_skip(BitReader& b, int n) { b.buf >>= n; b.cnt -= n; }
AMP_WD unsigned br_bits(BitReader& b, int n) { br_refill(b); const unsigned v = br_peek(b, n); br_skip(b, n); return v; }

AMP_WD uint8_t ld_cg_u8(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
AMP_WD unsigned rev_bits(unsigned v, int n) {       // the low n bits of v in reverse order
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - n);
#else
    unsigned r = 0;
    for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}

// canonical lists from code lengths (lane 0); returns false for an over-subscribed or incomplete code (a single-code distance
// tree is allowed, as zlib allows it)
AMP_WD bool huff_lists(HuffLists& h, const uint8_t* len, int n) {
    for (int l = 0; l < 16; ++l) h.count[l] = 0;
    for (int s = 0; s < n; ++s) ++h.count[len[s]];
    if (h.count[0] == n) return true;               // no codes at all: legal as long as none is used
    int left = 1;
    for (int l = 1; l < 16; ++l) { left <<= 1; left -= h.count[l]; if (left < 0) return false; }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + h.count[l]);
    for (int s = 0; s < n; ++s) if (len[s]) h.symbol[offs[len[s]]++] = (uint16_t)s;
    return left == 0 || (n - h.count[0] == 1);
}
// direct table from the lists, all lanes: entry of sorted position i = code first[len] + (i - offset[len]), bit-reversed, at stride 2^len
AMP_WD void huff_table(uint16_t* tab, int tbits, const HuffLists& h, int lane) {
    for (int i = lane; i < (1 << tbits); i += 32) tab[i] = 0;
    w_sync();
    int total = 0;
    for (int l = 1; l < 16; ++l) total += h.count[l];
    for (int i = lane; i < total; i += 32) {
        int l = 1, off = 0, first = 0;
        while (i >= off + h.count[l]) { off += h.count[l]; first = (first + h.count[l]) << 1; ++l; }
        if (l > tbits) continue;
        const unsigned code = (unsigned)(first + (i - off));
        const uint16_t e = (uint16_t)((h.symbol[i] << 4) | l);
        for (unsigned r = rev_bits(code, l); r < (1u << tbits); r += 1u << l) tab[r] = e;
    }
    w_sync();
}
// one symbol (lane 0): direct table, else the canonical walk bit by bit; -1 = invalid code
AMP_WD int huff_decode(BitReader& b, const uint16_t* tab, int tbits, const HuffLists& h) {
    br_refill(b);
    const unsigned e = tab[br_peek(b, tbits)];
    if (e) { br_skip(b, (int)(e & 15u)); return (int)(e >> 4); }
    int code = 0, first = 0, index = 0;
    unsigned long long bits = b.buf;
    for (int l = 1; l < 16; ++l) {
        code |= (int)(bits & 1ULL); bits >>= 1;
        const int c = h.count[l];
        if (code - c < first) { br_skip(b, l); return h.symbol[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

// Inflate one raw deflate stream of in_len bytes into out[0, out_len) (one warp; every lane calls).  Returns AMPZ_E_* bits.
// `in` must be readable up to the next 4-byte boundary past its end; `out` is written only inside [0, out_len).
// Output bytes go to the circular window in shared memory first and reach `out` in stretches copied by all lanes; a match whose
// source still lies in the window is copied inside shared memory, a farther one is read back from `out` (through L2).
AMP_WD int inflate_block(const uint8_t* in, long long in_len, uint8_t* out, long long out_len, InflateMem& M, int lane) {
    BitReader b; b.words = nullptr; b.ring = M.ring; b.nwords = 0; b.widx = 0; b.ring_hi = 0; b.buf = 0; b.cnt = 0;
    br_init(b, in, in + in_len, M.ring);            // every lane keeps the geometry; only lane 0's reader advances
    int o = 0;                                       // bytes produced (lane 0's copy is authoritative, broadcast where the lanes meet)
    int flushed = 0;                                 // bytes of the window already copied to `out`
    int wlo = 0;                                     // the window is valid for positions >= wlo (and >= o - AMPZ_WIN)
    int err = 0;
    // keep the ring ahead of the reader / copy finished output: all lanes, at points where they meet (widx, o are uniform then)
    auto top_up = [&](int widx) {
        while (b.ring_hi < b.nwords && b.ring_hi - widx < AMPZ_RING - 32) {
            const int idx = b.ring_hi + lane;
            if (idx < b.nwords) M.ring[idx & (AMPZ_RING - 1)] = b.words[idx];
            b.ring_hi += 32;
        }
        w_sync();
    };
    auto flush_to = [&](int upto) {
        for (int p = flushed + lane; p < upto; p += 32) out[p] = M.win[p & (AMPZ_WIN - 1)];
        flushed = upto;
        w_sync();
    };
    top_up(1);
    for (;;) {
        // ---- block header (lane 0), tables (all lanes)
        int last = 0, type = 0;
        if (lane == 0) { last = (int)br_bits(b, 1); type = (int)br_bits(b, 2); }
        last = w_shfl(last, 0); type = w_shfl(type, 0);
        if (type == 3) { err |= AMPZ_E_DATA; break; }
        if (type == 0) {                             // stored: LEN, NLEN, bytes (lane 0 copies; rare in BAM files)
            flush_to(o);
            int n = -1;
            if (lane == 0) {
                br_skip(b, b.cnt & 7);               // to the byte boundary
                const unsigned len = br_bits(b, 16), nlen = br_bits(b, 16);
                if ((len ^ nlen) == 0xFFFFu && (long long)o + (long long)len <= out_len) {
                    n = (int)len;
                    for (int i = 0; i < n; ++i) out[o + i] = (uint8_t)br_bits(b, 8);
                }
            }
            n = w_shfl(n, 0);
            if (n < 0) { err |= AMPZ_E_DATA; break; }
            o += n; flushed = o; wlo = o;            // these bytes are in `out` only
            top_up(w_shfl(b.widx, 0));
            if (last) break;
            continue;
        }
        int ok = 1;
        if (type == 1) {                             // fixed code (RFC 1951 3.2.6)
            for (int s = lane; s < 288; s += 32) M.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
            w_sync();
            if (lane == 0) ok = huff_lists(M.lit, M.lens, 288);
            w_sync();
            for (int s = lane; s < 30; s += 32) M.lens[s] = 5;
            w_sync();
            if (lane == 0) { huff_lists(M.dist, M.lens, 30); ok = 1; }      // (30 of the 32 five-bit codes: incomplete by definition)
        } else if (lane == 0) {                      // dynamic code: the code-length code first, then both length lists
            const int nlen = (int)br_bits(b, 5) + 257, ndist = (int)br_bits(b, 5) + 1, ncode = (int)br_bits(b, 4) + 4;
            if (nlen > 286 || ndist > 30) ok = 0;
            for (int i = 0; i < 19; ++i) M.lens[i] = 0;
            for (int i = 0; i < ncode; ++i) M.lens[kClOrder[i]] = (uint8_t)br_bits(b, 3);
            if (ok) ok = huff_lists(M.lit, M.lens, 19) ? 1 : 0;        // (the literal list's storage doubles as the code-length code's)
            if (ok) {
                // the code-length code is decoded with the canonical walk only (no direct table: at most 316 symbols)
                uint8_t* L = M.lens;                                  // overwritten in place once the 19 lengths are in the lists
                int idx = 0;
                HuffLists& cl = M.lit;
                // copy the code-length lists aside: M.lit is rebuilt below
                uint16_t ccount[16], csym[19];
                for (int i = 0; i < 16; ++i) ccount[i] = cl.count[i];
                for (int i = 0; i < 19; ++i) csym[i] = cl.symbol[i];
                while (idx < nlen + ndist && ok) {
                    br_refill(b);
                    int code = 0, first = 0, index = 0, sym = -1;
                    unsigned long long bits = b.buf;
                    for (int l = 1; l < 8; ++l) {
                        code |= (int)(bits & 1ULL); bits >>= 1;
                        const int c = ccount[l];
                        if (code - c < first) { br_skip(b, l); sym = csym[index + (code - first)]; break; }
                        index += c; first += c; first <<= 1; code <<= 1;
                    }
                    if (sym < 0) { ok = 0; break; }
                    if (sym < 16) L[idx++] = (uint8_t)sym;
                    else {
                        int rep, val = 0;
                        if (sym == 16) { if (idx == 0) { ok = 0; break; } val = L[idx - 1]; rep = 3 + (int)br_bits(b, 2); }
                        else if (sym == 17) rep = 3 + (int)br_bits(b, 3);
                        else rep = 11 + (int)br_bits(b, 7);
                        if (idx + rep > nlen + ndist) { ok = 0; break; }
                        while (rep--) L[idx++] = (uint8_t)val;
                    }
                }
                if (ok && L[256] == 0) ok = 0;                         // no end-of-block code
                if (ok) {
                    uint8_t dl[30];
                    for (int i = 0; i < ndist; ++i) dl[i] = L[nlen + i];
                    ok = huff_lists(M.lit, L, nlen) ? 1 : 0;
                    if (ok) ok = huff_lists(M.dist, dl, ndist) ? 1 : 0;
                }
            }
        }
        ok = w_shfl(ok, 0);
        if (!ok) { err |= AMPZ_E_DATA; break; }
        w_sync();
        huff_table(M.ltab, AMPZ_LBITS, M.lit, lane);
        huff_table(M.dtab, AMPZ_DBITS, M.dist, lane);
        top_up(w_shfl(b.widx, 0));
        // ---- symbols: lane 0 puts literals into the window as they come and stops at every match / end of block / error, and
        // when the ring runs low or the window fills up (len = -2: the lanes only top up / flush and it goes on)
        for (;;) {
            int len = 0, dist = 0;                   // len > 0: match; len == 0: end of block; len == -1: error; -2: service stop
            if (lane == 0) {
                for (;;) {
                    if ((b.ring_hi < b.nwords && b.ring_hi - b.widx < 4) || o - flushed > AMPZ_WIN - 320) { len = -2; break; }
                    const int sym = huff_decode(b, M.ltab, AMPZ_LBITS, M.lit);
                    if (sym < 0) { len = -1; break; }
                    if (sym < 256) {
                        if ((long long)o >= out_len) { len = -1; break; }
                        M.win[o & (AMPZ_WIN - 1)] = (uint8_t)sym; ++o;
                        continue;
                    }
                    if (sym == 256) break;
                    if (sym > 285) { len = -1; break; }
                    len = kLenBase[sym - 257] + (int)br_bits(b, kLenExtra[sym - 257]);
                    const int ds = huff_decode(b, M.dtab, AMPZ_DBITS, M.dist);
                    if (ds < 0 || ds > 29) { len = -1; break; }
                    dist = kDistBase[ds] + (int)br_bits(b, kDistExtra[ds]);
                    if (dist > o || (long long)o + len > out_len) len = -1;
                    break;
                }
            }
            {   // length and distance in one word
                const int ld = w_shfl(len > 0 ? (len << 16) | dist : len, 0);
                len = ld > 0 ? ld >> 16 : ld; dist = ld & 0xFFFF;
            }
            if (len == -1) { err |= AMPZ_E_DATA; break; }
            o = w_shfl(o, 0);
            const int widx = w_shfl(b.widx, 0);
            if (len == 0) { top_up(widx); break; }
            if (len > 0) {
                // window[o + i] = byte at o - dist + (i mod dist): the stretch [o - dist, o) is complete (periodic extension)
                const int s0 = o - dist;
                if (s0 >= wlo && dist <= AMPZ_WIN - 258) {
                    if (dist >= len) { for (int i = lane; i < len; i += 32) M.win[(o + i) & (AMPZ_WIN - 1)] = M.win[(s0 + i) & (AMPZ_WIN - 1)]; }
                    else if (dist == 1) { const uint8_t v = M.win[s0 & (AMPZ_WIN - 1)]; for (int i = lane; i < len; i += 32) M.win[(o + i) & (AMPZ_WIN - 1)] = v; }
                    else for (int i = lane; i < len; i += 32) M.win[(o + i) & (AMPZ_WIN - 1)] = M.win[(s0 + i % dist) & (AMPZ_WIN - 1)];
                } else {                             // far source: everything before o is in `out` after the flush (read through L2)
                    flush_to(o);
                    for (int i = lane; i < len; i += 32) M.win[(o + i) & (AMPZ_WIN - 1)] = ld_cg_u8(out + s0 + (dist >= len ? i : i % dist));
                }
                w_sync();
                o += len;
            }
            if (o - flushed >= 512) flush_to(o);
            top_up(widx);
        }
        if (err || last) break;
    }
    o = w_shfl(o, 0);
    err = w_shfl(err, 0) | err;
    if (!err) flush_to(o);
    if (!err && (long long)o != out_len) err |= AMPZ_E_SIZE;
    return err;
}

// ---- BAM record chains -------------------------------------------------------------------------------------------------------
// record (after the 4-byte block_size): refID, pos, l_read_name(u8), mapq(u8), bin(u16), n_cigar(u16), flag(u16), l_seq, next_refID,
// next_pos, tlen, read_name, cigar, seq, qual, tags
AMP_HD uint32_t ld_u32u(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
AMP_HD uint32_t ld_u16u(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

struct BamBlockTotals { unsigned int n_rec, n_cig; unsigned long long n_seq, n_qual; };   // per BGZF block

// one thread: records of raw[lo, hi); false when the chain does not end exactly at hi or a header does not fit its record
AMP_HD bool bam_chain_totals(const uint8_t* raw, long long lo, long long hi, BamBlockTotals& t) {
    t.n_rec = 0; t.n_cig = 0; t.n_seq = 0; t.n_qual = 0;
    long long p = lo;
    while (p < hi) {
        if (p + 36 > hi) return false;
        const long long bs = ld_u32u(raw + p);
        const uint8_t* r = raw + p + 4;
        const long long lname = r[8], nc = ld_u16u(r + 12), ls = (int)ld_u32u(r + 16);
        if (bs < 32 || p + 4 + bs > hi || ls < 0 || 32 + lname + 4 * nc + (ls + 1) / 2 + ls > bs) return false;
        ++t.n_rec; t.n_cig += (unsigned int)nc; t.n_seq += (unsigned long long)((ls + 1) / 2); t.n_qual += (unsigned long long)ls;
        p += 4 + bs;
    }
    return p == hi;
}

struct BamSoa {       // destination arrays (device), indexed with global read indices / offsets
    int32_t* pos; uint16_t* flag; int32_t* tlen; uint32_t* cig_off; uint32_t* cigar; uint32_t* seq_off; uint8_t* seq;
    uint32_t* qual_off; uint8_t* qual; unsigned long long* rec_off;   // rec_off: offset of the record's block_size word in the raw stream
};
// one warp: the records of raw[lo, hi) into the arrays, starting at read r0 / CIGAR op c0 / seq byte s0 / qual byte q0
AMP_WD void bam_scatter_block(const uint8_t* raw, long long lo, long long hi, const BamSoa& D, unsigned long long r0, unsigned long long c0,
                              unsigned long long s0, unsigned long long q0, int lane) {
    long long p = lo;
    unsigned long long ri = r0, ci = c0, si = s0, qi = q0;
    while (p < hi) {
        const uint8_t* r = raw + p + 4;
        const long long bs = ld_u32u(raw + p);
        const uint32_t lname = r[8], nc = ld_u16u(r + 12), ls = ld_u32u(r + 16);
        if (lane == 0) {
            D.pos[ri] = (int32_t)ld_u32u(r + 4); D.flag[ri] = (uint16_t)ld_u16u(r + 14); D.tlen[ri] = (int32_t)ld_u32u(r + 28);
            D.cig_off[ri] = (uint32_t)ci; D.seq_off[ri] = (uint32_t)si; D.qual_off[ri] = (uint32_t)qi;
            D.rec_off[ri] = (unsigned long long)p;
        }
        const uint8_t* c = r + 32 + lname;
        for (uint32_t k = lane; k < nc; k += 32) D.cigar[ci + k] = ld_u32u(c + 4 * k);
        const uint8_t* s = c + 4 * (size_t)nc;
        const uint32_t nsb = (ls + 1) / 2;
        for (uint32_t k = lane; k < nsb; k += 32) D.seq[si + k] = s[k];
        const uint8_t* q = s + nsb;
        for (uint32_t k = lane; k < ls; k += 32) D.qual[qi + k] = q[k];
        ++ri; ci += nc; si += nsb; qi += ls;
        p += 4 + bs;
    }
}

// ---- trimmed BAM records rebuilt from the inflated input (what assigning cigartuples / reference_start does to a pysam segment
// before out_aln.write, AmpliPy.py:463-514, 591-658, 911): everything but block_size / pos / bin / n_cigar / CIGAR byte for byte ----
AMP_HD int bam_reg2bin(long long beg, long long end) {   // SAM spec 5.3
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}
// size of the record at raw + rec (its block_size word) once its CIGAR has nc_new ops
AMP_HD uint32_t bam_new_record_size(const uint8_t* raw, unsigned long long rec, uint32_t nc_new) {
    const uint8_t* r = raw + rec;
    return 4u + ld_u32u(r) - 4u * ld_u16u(r + 4 + 12) + 4u * nc_new;
}
// one warp: the record at raw + rec with position new_pos and CIGAR cg[0, nc_new) written to w
AMP_WD void bam_rewrite_record(const uint8_t* raw, unsigned long long rec, int32_t new_pos, const uint32_t* cg, uint32_t nc_new, uint8_t* w, int lane) {
    const uint8_t* r = raw + rec + 4;
    const uint32_t bs = ld_u32u(r - 4), lname = r[8], nc_old = ld_u16u(r + 12), flag = ld_u16u(r + 14);
    const uint32_t nbs = bs - 4u * nc_old + 4u * nc_new;
    const uint32_t head = 32u + lname, rest = bs - head - 4u * nc_old;
    const uint8_t* tail = r + head + 4u * nc_old;
    uint8_t* wt = w + 4 + head + 4u * nc_new;
    for (uint32_t k = lane; k < head; k += 32) w[4 + k] = r[k];
    for (uint32_t k = lane; k < rest; k += 32) wt[k] = tail[k];
    for (uint32_t k = lane; k < 4u * nc_new; k += 32) w[4 + head + k] = (uint8_t)(cg[k >> 2] >> (8 * (k & 3)));
    w_sync();
    if (lane == 0) {
        long long rlen = 0;
        for (uint32_t c = 0; c < nc_new; ++c) { const uint32_t op = cg[c] & 15u; if ((0x18Du >> op) & 1u) rlen += cg[c] >> 4; }
        if ((flag & 4u) || rlen == 0) rlen = 1;                                // htslib bam_endpos
        const uint32_t bin = (uint32_t)bam_reg2bin(new_pos, new_pos + rlen);
        for (int b = 0; b < 4; ++b) { w[b] = (uint8_t)(nbs >> (8 * b)); w[8 + b] = (uint8_t)((uint32_t)new_pos >> (8 * b)); }
        w[14] = (uint8_t)bin; w[15] = (uint8_t)(bin >> 8); w[16] = (uint8_t)nc_new; w[17] = (uint8_t)(nc_new >> 8);
    }
}

}  // namespace amp
