// amp_hostio.cpp -- host-side BGZF/BAM codec feeding the C ABI (libamplipy_hostio.so, CPU only).
//
// Replaces what the reference gets from pysam/htslib around its per-read loop
// (/root/reference/AmpliPy.py:296-360 create_AlignmentFile_objects, 896 iteration, 911 out_aln.write):
//   * multi-threaded BGZF inflate / deflate (independent 64 KiB blocks)
//   * BAM record scan -> flat struct-of-arrays buffers (pos, flag, tlen, packed CIGAR, 4-bit seq, qual)
//   * BAM record rebuild with patched pos / bin / n_cigar / CIGAR (what assigning `cigartuples` and
//     `reference_start` does to a pysam segment); everything else is copied byte for byte.
// Build: g++ -O2 -fPIC -shared -fopenmp amp_hostio.cpp -lz
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

// ---- BGZF -------------------------------------------------------------------------------------------
// Scan block boundaries.  Returns the number of blocks (or -1 on a malformed stream); fills
// in_off[k] (start of block k), out_len[k] (ISIZE).  Call with max_blocks = 0 to count only.
long long amp_bgzf_scan(const uint8_t* in, long long in_len, long long* in_off, uint32_t* out_len, long long max_blocks) {
    long long p = 0, k = 0;
    while (p + 18 <= in_len) {
        if (in[p] != 31 || in[p + 1] != 139 || in[p + 2] != 8 || !(in[p + 3] & 4)) return -1;
        const uint32_t xlen = rd16(in + p + 10);
        long long x = p + 12, xend = x + xlen;
        if (xend > in_len) return -1;                                   // extra field runs past the buffer
        long long bsize = -1;
        while (x + 4 <= xend) {
            const uint32_t slen = rd16(in + x + 2);
            if (x + 4 + (long long)slen > xend) return -1;              // subfield runs past the extra field
            if (in[x] == 66 && in[x + 1] == 67 && slen == 2) bsize = (long long)rd16(in + x + 4) + 1;
            x += 4 + slen;
        }
        // header (12) + extra field + trailer (CRC32, ISIZE) must fit inside the block, the block inside the buffer
        if (bsize < 0 || 12 + (long long)xlen + 8 > bsize || p + bsize > in_len) return -1;
        if (max_blocks) {
            if (k >= max_blocks) return -1;
            in_off[k] = p; out_len[k] = rd32(in + p + bsize - 4);
        }
        ++k; p += bsize;
    }
    return p == in_len ? k : -1;
}

// Inflate all blocks in parallel into out (out_off = exclusive prefix sum of out_len).  0 on success.
int amp_bgzf_inflate(const uint8_t* in, const long long* in_off, const uint32_t* out_len, const long long* out_off,
                     long long n_blocks, long long in_len, uint8_t* out, int n_threads) {
    int bad = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(| : bad)
    for (long long k = 0; k < n_blocks; ++k) {
        const uint8_t* b = in + in_off[k];
        const long long bend = (k + 1 < n_blocks) ? in_off[k + 1] : in_len;
        const uint32_t xlen = rd16(b + 10);
        const uint8_t* cdata = b + 12 + xlen;
        const long long clen = bend - in_off[k] - 12 - xlen - 8;
        if (clen < 0 || clen > 0x7fffffffLL) { bad |= 1; continue; }
        if (out_len[k] == 0) continue;
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) { bad |= 1; continue; }
        zs.next_in = (Bytef*)cdata; zs.avail_in = (uInt)clen;
        zs.next_out = out + out_off[k]; zs.avail_out = out_len[k];
        const int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.total_out != out_len[k]) bad |= 1;
        inflateEnd(&zs);
    }
    return bad;
}

// Deflate `len` bytes into BGZF blocks of <= 0xff00 payload bytes each, in parallel, plus the EOF block.
// out capacity must be >= amp_bgzf_bound(len).  Returns bytes written or -1.
long long amp_bgzf_bound(long long len) { return (len / 0xff00 + 2) * (0xff00 + 1024) + 28; }
// Deflate in[bstart[k], bstart[k+1]) as BGZF block k (each at most 0xff00 bytes), in parallel, plus the EOF block.
long long amp_bgzf_deflate_blocks(const uint8_t* in, const long long* bstart, long long nb, uint8_t* out, int level, int n_threads) {
    const long long BS = 0xff00;
    const long long slot = BS + 1024;
    uint32_t* clen = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(nb + 1));
    uint8_t* tmp = (uint8_t*)malloc((size_t)(nb ? nb : 1) * (size_t)slot);
    int bad = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 8) reduction(| : bad)
    for (long long k = 0; k < nb; ++k) {
        const uint8_t* src = in + bstart[k];
        const long long n64 = bstart[k + 1] - bstart[k];
        if (n64 < 0 || n64 > BS) { bad |= 1; continue; }
        const uInt n = (uInt)n64;
        uint8_t* dst = tmp + k * slot;
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad |= 1; continue; }
        zs.next_in = (Bytef*)src; zs.avail_in = n; zs.next_out = dst + 18; zs.avail_out = (uInt)(slot - 26);
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) bad |= 1;
        const uint32_t c = (uint32_t)zs.total_out;
        deflateEnd(&zs);
        const uint32_t bsize = c + 26;
        if (bsize > 65536) { bad |= 1; continue; }
        static const uint8_t hdr[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
        memcpy(dst, hdr, 12); dst[12] = 66; dst[13] = 67; dst[14] = 2; dst[15] = 0;
        const uint16_t bs1 = (uint16_t)(bsize - 1); memcpy(dst + 16, &bs1, 2);
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, n);
        memcpy(dst + 18 + c, &crc, 4); const uint32_t isz = n; memcpy(dst + 22 + c, &isz, 4);
        clen[k] = bsize;
    }
    long long o = 0;
    if (!bad) {
        for (long long k = 0; k < nb; ++k) { memcpy(out + o, tmp + k * slot, clen[k]); o += clen[k]; }
        static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        memcpy(out + o, eof, 28); o += 28;
    }
    free(clen); free(tmp);
    return bad ? -1 : o;
}
long long amp_bgzf_deflate(const uint8_t* in, long long len, uint8_t* out, int level, int n_threads) {
    const long long BS = 0xff00;
    const long long nb = (len + BS - 1) / BS;
    long long* bstart = (long long*)malloc(sizeof(long long) * (size_t)(nb + 1));
    for (long long k = 0; k <= nb; ++k) bstart[k] = k * BS < len ? k * BS : len;
    const long long r = amp_bgzf_deflate_blocks(in, bstart, nb, out, level, n_threads);
    free(bstart);
    return r;
}
// Block starts the way htslib lays a BAM out (bgzf_flush after the header, bgzf_flush_try per record): a block holds whole
// units -- bounds[0..n_bounds) are the cut points (sorted, bounds[0] = 0, the last one = the length) -- up to 0xff00 bytes; a
// unit longer than that is split.  Returns the number of blocks (bstart gets one entry more); call with bstart = NULL to count.
long long amp_bgzf_plan(const long long* bounds, long long n_bounds, long long* bstart, long long max_blocks) {
    const long long BS = 0xff00;
    long long nb = 0, cur = 0;                                  // cur = start of the open block
    if (n_bounds < 2) { if (bstart && max_blocks >= 0) bstart[0] = 0; return 0; }
    for (long long i = 1; i < n_bounds; ++i) {
        const long long end = bounds[i];
        if (end - cur <= BS) continue;                          // unit i fits the open block
        if (bounds[i - 1] > cur) {                              // close the open block in front of unit i
            if (bstart) { if (nb >= max_blocks) return -1; bstart[nb] = cur; }
            ++nb; cur = bounds[i - 1];
        }
        while (end - cur > BS) {                                // a unit longer than a block
            if (bstart) { if (nb >= max_blocks) return -1; bstart[nb] = cur; }
            ++nb; cur += BS;
        }
    }
    if (bounds[n_bounds - 1] > cur) {
        if (bstart) { if (nb >= max_blocks) return -1; bstart[nb] = cur; }
        ++nb;
    }
    if (bstart) bstart[nb] = bounds[n_bounds - 1];
    return nb;
}

// ---- BAM records -> struct of arrays ------------------------------------------------------------------
// Record layout (after the 4-byte block_size): refID, pos, l_read_name(u8), mapq(u8), bin(u16), n_cigar(u16),
// flag(u16), l_seq, next_refID, next_pos, tlen, read_name, cigar, seq, qual, tags.
long long amp_bam_scan(const uint8_t* buf, long long len, long long start, long long* rec_off, int32_t* n_cigar, int32_t* l_seq,
                       long long max_records) {
    long long p = start, k = 0;
    while (p + 4 <= len) {
        const uint32_t bs = rd32(buf + p);
        if (bs < 32 || p + 4 + (long long)bs > len) return -1;
        {   // the variable-length fields named by the fixed header must fit inside the record (untrusted input)
            const uint8_t* r = buf + p + 4;
            const long long lname = r[8], nc = rd16(r + 12), ls = (int32_t)rd32(r + 16);
            if (ls < 0 || 32 + lname + 4 * nc + (ls + 1) / 2 + ls > (long long)bs) return -1;
        }
        if (max_records) {
            if (k >= max_records) return -1;
            rec_off[k] = p; n_cigar[k] = rd16(buf + p + 4 + 12); l_seq[k] = (int32_t)rd32(buf + p + 4 + 16);
        }
        ++k; p += 4 + bs;
    }
    return p == len ? k : -1;
}

// qualities stored as 0xFF (absent) are passed through as-is; the reference cannot process such reads either
void amp_bam_fill(const uint8_t* buf, const long long* rec_off, long long n, int32_t* pos, uint16_t* flag, int32_t* tlen,
                  const uint32_t* cig_off, uint32_t* cigar, const uint32_t* seq_off, uint8_t* seq, const uint32_t* qual_off,
                  uint8_t* qual, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) {
        const uint8_t* r = buf + rec_off[i] + 4;
        pos[i] = (int32_t)rd32(r + 4);
        const uint32_t lname = r[8];
        const uint32_t nc = rd16(r + 12);
        flag[i] = rd16(r + 14);
        const uint32_t ls = rd32(r + 16);
        tlen[i] = (int32_t)rd32(r + 28);
        const uint8_t* p = r + 32 + lname;
        memcpy(cigar + cig_off[i], p, 4 * (size_t)nc); p += 4 * (size_t)nc;
        memcpy(seq + seq_off[i], p, (ls + 1) / 2); p += (ls + 1) / 2;
        memcpy(qual + qual_off[i], p, ls);
    }
}

static inline int reg2bin(int64_t beg, int64_t end);
// A ReadBatch as BAM records (names from a blob of NUL-terminated strings, or r<i>; refID 0, mates on the same reference):
// synthetic inputs for tests and benchmarks.
// rec_off gets n + 1 offsets; pass out = NULL to get the size.
long long amp_bam_serialize(long long n, const int32_t* pos, const uint16_t* flag, const int32_t* tlen, const uint32_t* cig_off,
                            const uint32_t* cigar, const uint32_t* seq_off, const uint8_t* seq, const uint32_t* qual_off,
                            const uint8_t* qual, int mapq, const char* names, const long long* name_off, uint8_t* out, long long* rec_off) {
    long long o = 0;
    for (long long i = 0; i < n; ++i) {
        char nbuf[32];
        const char* name = nbuf;
        int lname;
        if (names) { name = names + name_off[i]; lname = (int)(name_off[i + 1] - name_off[i]); }     // NUL-terminated, <= 255 bytes
        else lname = snprintf(nbuf, sizeof nbuf, "r%lld", i) + 1;
        const uint32_t nc = cig_off[i + 1] - cig_off[i], ls = qual_off[i + 1] - qual_off[i], nsb = seq_off[i + 1] - seq_off[i];
        const uint32_t bs = 32 + (uint32_t)lname + 4 * nc + nsb + ls;
        if (rec_off) rec_off[i] = o;
        if (out) {
            uint8_t* w = out + o;
            memcpy(w, &bs, 4);
            const uint32_t* cg = cigar + cig_off[i];
            int64_t rlen = 0;
            for (uint32_t c = 0; c < nc; ++c) { const uint32_t op = cg[c] & 15; if ((0x18Du >> op) & 1) rlen += cg[c] >> 4; }
            const uint16_t fl = flag[i];
            if ((fl & 4) || rlen == 0) rlen = 1;
            const int32_t rid = (fl & 4) ? -1 : 0, p = pos[i], nrid = (fl & 1) ? 0 : -1;
            int32_t npos = (fl & 1) ? p + tlen[i] : -1; if ((fl & 1) && npos < 0) npos = 0;
            const uint16_t bin = (uint16_t)reg2bin(p, p + rlen), ncw = (uint16_t)nc;
            const uint8_t ln = (uint8_t)lname, mq = (uint8_t)mapq;
            const int32_t lsi = (int32_t)ls, tl = tlen[i];
            memcpy(w + 4, &rid, 4); memcpy(w + 8, &p, 4); w[12] = ln; w[13] = mq; memcpy(w + 14, &bin, 2); memcpy(w + 16, &ncw, 2);
            memcpy(w + 18, &fl, 2); memcpy(w + 20, &lsi, 4); memcpy(w + 24, &nrid, 4); memcpy(w + 28, &npos, 4); memcpy(w + 32, &tl, 4);
            memcpy(w + 36, name, (size_t)lname);
            uint8_t* q = w + 36 + lname;
            memcpy(q, cg, 4 * (size_t)nc); q += 4 * (size_t)nc;
            memcpy(q, seq + seq_off[i], nsb); q += nsb;
            memcpy(q, qual + qual_off[i], ls);
        }
        o += 4 + bs;
    }
    if (rec_off) rec_off[n] = o;
    return o;
}

static inline int reg2bin(int64_t beg, int64_t end) {   // SAM spec 5.3
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

// Rebuild the selected records with new pos / CIGAR.  sel[k] = index of the k-th output record; new CIGAR of
// read i = new_cigar[cig_off[i] + 3*i .. + new_ncig[i]).  Pass out = NULL to get the required size.
long long amp_bam_rewrite(const uint8_t* buf, const long long* rec_off, const long long* sel, long long n_sel,
                          const int32_t* new_pos, const uint16_t* new_ncig, const uint32_t* cig_off, const uint32_t* new_cigar,
                          uint8_t* out) {
    // sizes first (one pass), then the records in parallel at their offsets
    long long* ooff = (long long*)malloc(sizeof(long long) * (size_t)(n_sel + 1));
    long long o = 0;
    for (long long k = 0; k < n_sel; ++k) {
        const long long i = sel[k];
        const uint8_t* r = buf + rec_off[i] + 4;
        const uint32_t bs = rd32(buf + rec_off[i]);
        const uint32_t nc_old = rd16(r + 12), nc_new = new_ncig[i];
        ooff[k] = o;
        o += 4 + (long long)(bs - 4 * nc_old + 4 * nc_new);
    }
    ooff[n_sel] = o;
    if (out) {
#pragma omp parallel for schedule(static)
        for (long long k = 0; k < n_sel; ++k) {
            const long long i = sel[k];
            const uint8_t* r = buf + rec_off[i] + 4;
            const uint32_t bs = rd32(buf + rec_off[i]);
            const uint32_t lname = r[8], nc_old = rd16(r + 12), nc_new = new_ncig[i];
            const uint32_t nbs = bs - 4 * nc_old + 4 * nc_new;
            uint8_t* w = out + ooff[k];
            memcpy(w, &nbs, 4);
            memcpy(w + 4, r, 32 + lname);
            const int32_t p = new_pos[i]; memcpy(w + 4 + 4, &p, 4);
            const uint32_t* cg = new_cigar + (size_t)cig_off[i] + 3 * (size_t)i;
            int64_t rlen = 0;
            for (uint32_t c = 0; c < nc_new; ++c) { const uint32_t op = cg[c] & 15; if ((0x18Du >> op) & 1) rlen += cg[c] >> 4; }
            const uint16_t flag = rd16(r + 14);
            if ((flag & 4) || rlen == 0) rlen = 1;                       // htslib bam_endpos
            const uint16_t bin = (uint16_t)reg2bin(p, p + rlen); memcpy(w + 4 + 10, &bin, 2);
            const uint16_t ncw = (uint16_t)nc_new; memcpy(w + 4 + 12, &ncw, 2);
            memcpy(w + 4 + 32 + lname, cg, 4 * (size_t)nc_new);
            const uint32_t rest = bs - 32 - lname - 4 * nc_old;
            memcpy(w + 4 + 32 + lname + 4 * (size_t)nc_new, r + 32 + lname + 4 * (size_t)nc_old, rest);
        }
    }
    free(ooff);
    return o;
}
// offsets of the records written by amp_bam_rewrite (n_sel + 1 entries): the block planner's cut points
void amp_bam_rewrite_offsets(const uint8_t* buf, const long long* rec_off, const long long* sel, long long n_sel,
                             const uint16_t* new_ncig, long long* ooff) {
    long long o = 0;
    for (long long k = 0; k < n_sel; ++k) {
        const long long i = sel[k];
        const uint32_t bs = rd32(buf + rec_off[i]);
        const uint32_t nc_old = rd16(buf + rec_off[i] + 4 + 12), nc_new = new_ncig[i];
        ooff[k] = o;
        o += 4 + (long long)(bs - 4 * nc_old + 4 * nc_new);
    }
    ooff[n_sel] = o;
}

int amp_hostio_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
