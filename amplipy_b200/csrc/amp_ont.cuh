// amp_ont.cuh -- fused trim + pileup for indel-rich batches (ONT-like: ~400 bases, tens of CIGAR ops per read), sm_100a.
//
// One warp per read, the lanes over the read's CIGAR ops.  trim_read (AmpliPy.py:426-687) never reorders ops: each of its three
// steps only turns ops at one end of the alignment into soft clip (or drops them), so the trimmed CIGAR is
//     [H] [S lead]  ops[ka .. kb] (first one shortened by fc bases at its front, last one kept for bk bases)  [S trail] [H]
// and the three steps reduce to finding ka / kb on the prefix sums of the query / reference lengths of the ops:
//   start clip   (450-514)  T = max_primer_end[pos] + 1: the first reference-consuming op that reaches T is cut there (get_pos_on_query
//                389-412 uses <=, so a target on an op's end falls to the boundary case); on a boundary the alignment restarts at the
//                next M/=/X: insertions in between turn into soft clip (487-488), deletions are dropped and advance the start
//   end clip     (516-558)  the same from the other side with R = min_primer_start[reference_end - 1]
//   quality clip (560-686)  del_len aligned query bases leave from the end (forward strand) or from the start (reverse strand, pos
//                stays: F6); an op is cut wherever the count runs out, insertions included, and on a boundary the deletions right
//                behind it survive (qual_rewrite_op copies everything once del_len is 0)
// Everything the closed form does not cover (hard / soft clips in odd places, P ops, equal neighbours, reads that a clip swallows,
// offset-induced targets in front of the read, rows too long for the staging buffers) is appended to the CTA's list and takes the
// loop-for-loop generic path of amp_warp.cuh afterwards.
//
// update_base_counts (690-753): the aligned bases of the kept query range are dealt out evenly over the lanes (a lane finds the
// op of its first base by binary search on the prefix sums); then lane per op: the positions of a D/N op into the CTA's tile,
// and an I op resolves the insertion state machine (730-748) of its I op from the op itself and the kind of its successor
// (exit A: a match follows, B: a deletion follows -- the key runs to the end of the read, C: the trailing clip follows, D: a
// low-quality inserted base).  Insertion alleles are parked in the warp's own list and added to the table a lane each.
#pragma once
#include "amp_warp.cuh"

namespace amp {

#ifndef AMPO_WARPS
#define AMPO_WARPS 32
#endif
#ifndef AMPO_GWARPS
#define AMPO_GWARPS 2
#endif
#define AMPO_WT 1024
#ifdef AMPO_LOCKSTEP
#define AMPO_STEP_SYNC() c_sync()      // experiment: the CTA's warps in lock step (one code region at a time, but memory latency exposed)
#else
#define AMPO_STEP_SYNC() ((void)0)
#endif
#define AMPO_MAXOPS 128
#define AMPO_QCAP 512                         // staged quality bytes per read
#define AMPO_QBUF (AMP7_PAD + AMPO_QCAP + 16 + AMP7_QSLACK)
#define AMPO_SBUF (AMP7_PAD + AMPO_QCAP / 2 + 16 + AMP7_SSLACK)
#define AMPO_EVCAP 96                         // insertion alleles a warp parks before it adds them to the table, a lane each
#define AMPO_WARP_BYTES (AMPO_MAXOPS * 4 + 2 * (AMPO_MAXOPS + 8) * 4 + AMPO_QBUF + AMPO_SBUF + 16 + AMPO_EVCAP * 12 + 16)
struct OntMem { uint32_t* ops; int* q0; int* r0; uint8_t* qbuf; uint8_t* sbuf; unsigned long long* bar; int* ev; int* nev; };
AMP_HD size_t smem_bytes_ont(int wt, int warps, int gwarps) {
    return tile_bytes_v7(wt) + C7_WORDS * 4 + (size_t)warps * AMPO_WARP_BYTES + (size_t)gwarps * (AMP7_FAST_BYTES + AMP7_GEXTRA_BYTES);
}
AMP_HD OntMem carve_ont(unsigned char* base, int wt, int w) {
    unsigned char* b = base + tile_bytes_v7(wt) + C7_WORDS * 4 + (size_t)w * AMPO_WARP_BYTES;
    OntMem m;
    m.ops = (uint32_t*)b; b += AMPO_MAXOPS * 4;
    m.q0 = (int*)b; b += (AMPO_MAXOPS + 8) * 4;
    m.r0 = (int*)b; b += (AMPO_MAXOPS + 8) * 4;
    m.qbuf = b; b += AMPO_QBUF;
    m.sbuf = b; b += AMPO_SBUF;
    m.bar = (unsigned long long*)b; b += 16;
    m.ev = (int*)b; b += AMPO_EVCAP * 12;
    m.nev = (int*)b;
    return m;
}
// the generic phase's buffers of the g-th generic-capable warp (behind all per-warp blocks)
AMP_HD WarpMem7 carve_ont_generic(unsigned char* base, int wt, int warps, int g) {
    unsigned char* b = base + tile_bytes_v7(wt) + C7_WORDS * 4 + (size_t)warps * AMPO_WARP_BYTES + (size_t)g * (AMP7_FAST_BYTES + AMP7_GEXTRA_BYTES);
    WarpMem7 m;
    m.qbuf = b; b += AMP7_QBUF;
    m.sbuf = b; b += AMP7_SBUF;
    m.par = (Par4*)b; b += 32 * 32;
    m.own = (uint16_t*)b; b += 64;
    m.bar = (unsigned long long*)b; b += 16;
    m.runs = (Seg*)b; b += AMP7_RUNCAP * 16;
    m.queue = (uint32_t*)b; b += AMP7_QCAP * 4;
    m.ctr = (int*)b; b += 32;
    m.cig = (uint32_t*)b;
    m.ctrl = (int*)(base + tile_bytes_v7(wt));
    return m;
}

// first / last op index in [lo, hi] for which pred holds, -1 if none (warp-uniform; the lanes test 32 ops at a time)
template <class Pred>
AMP_WD int first_op(int lo, int hi, int lane, Pred pred) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int base = lo; base <= hi; base += 32) {
        const int k = base + lane;
        const unsigned m = w_ballot(k <= hi && pred(k));
        if (m) return base + ctz32(m);
    }
    return -1;
}
template <class Pred>
AMP_WD int last_op(int lo, int hi, int lane, Pred pred) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int top = hi; top >= lo; top -= 32) {
        const int k = top - 31 + lane;
        const unsigned m = w_ballot(k >= lo && pred(k));
        if (m) return top - 31 + msb32(m);
    }
    return -1;
}

// Sliding-window search over buf[a0, a0 + m), window 4, by the warp: the same result as window_del_len_fwd / _rev.  A lane tests
// the eight windows of one block (word-wise, as window_del_blocks does); the first (forward) / last (reverse) failing window
// decides, the three shrinking windows at the open end are checked from three bytes.  Reads up to 16 bytes past the run.
AMP_WD int warp_window_del(const uint8_t* buf, int a0, int m, bool rev, int minq, int lane) {
    if (m < 8) {
        int d = 0;
        if (lane == 0) d = rev ? window_del_len_rev(buf + a0, m, 4, minq) : window_del_len_fwd(buf + a0, m, 4, minq);
        return w_shfl(d, 0);
    }
    const uint32_t* A = (const uint32_t*)(buf + (a0 & ~3));
    const unsigned sh = (unsigned)(a0 & 3) << 3, nthr = (unsigned)(-4 * minq);
    const int nwin = m - 3, nb = (nwin + 7) >> 3;
    const unsigned last_mask = (1u << (nwin - 8 * (nb - 1))) - 1u;
    auto block_bits = [&](int b) -> unsigned {
        const unsigned x0 = A[2 * b], x1 = A[2 * b + 1], x2 = A[2 * b + 2], x3 = A[2 * b + 3];
        unsigned bits = win8_bits(funnel_r(x0, x1, sh), funnel_r(x1, x2, sh), funnel_r(x2, x3, sh), nthr);
        if (b == nb - 1) bits &= last_mask;
        return bits;
    };
    if (!rev) {
        for (int base = 0; base < nb; base += 32) {
            const int b = base + lane;
            const unsigned bits = b < nb ? block_bits(b) : 0u;
            const unsigned mm = w_ballot(bits != 0u);
            if (mm) { const int l = ctz32(mm); return m - (8 * (base + l) + ctz32((unsigned)w_shfl((int)bits, l))); }
        }
    } else {
        for (int top = nb - 1; top >= 0; top -= 32) {
            const int b = top - 31 + lane;
            const unsigned bits = b >= 0 ? block_bits(b) : 0u;
            const unsigned mm = w_ballot(bits != 0u);
            if (mm) { const int l = msb32(mm); return 8 * (top - 31 + l) + msb32((unsigned)w_shfl((int)bits, l)) + 4; }
        }
    }
    const uint8_t* e3 = buf + a0 + (rev ? 0 : m - 3);   // shrinking windows w = 3, 2, 1 at the open end
    const int x0 = e3[0], x1 = e3[1], x2 = e3[2];
    const int e = rev ? x0 : x2;
    if (x0 + x1 + x2 < 3 * minq) return 3;
    if (e + x1 < 2 * minq) return 2;
    if (e < minq) return 1;
    return 0;
}

// An insertion allele of the warp's read: parked in the warp's own list (no other warp touches it), added to the table by
// ont_drain with a lane per allele -- inside the per-op code an add would run with one or two lanes active, and on ONT-like data
// there are a dozen alleles per read.
AMP_WD void ont_defer(const KParams& P, const OntMem& wm, uint32_t so0, int pos, int b, int n) {
    if (b < 65536 && n > 0 && n < 65536) {
        const int idx = atomic_add(wm.nev, 1);
        if (idx < AMPO_EVCAP) { int* e = wm.ev + 3 * idx; e[0] = pos; e[1] = (int)so0; e[2] = (int)((uint32_t)b | ((uint32_t)n << 16)); return; }
    }
    ins_commit(P, P.b.seq + so0, pos, b, n);
}
AMP_WD_COLD void ont_drain(const KParams& P, const OntMem& wm, int lane) {
    w_sync();
    int n = *wm.nev; if (n > AMPO_EVCAP) n = AMPO_EVCAP;
    for (int k = lane; k < n; k += 32) {
        const int* e = wm.ev + 3 * k;
        ins_commit(P, P.b.seq + (uint32_t)e[1], e[0], (int)((uint32_t)e[2] & 0xFFFFu), (int)((uint32_t)e[2] >> 16));
    }
    w_sync();
    if (lane == 0) *wm.nev = 0;
    w_sync();
}

// One read by one warp, in four steps.  (AMPO_LOCKSTEP builds separate them by block barriers, so that the CTA's warps run the same
// stretch of code at any time: no instruction-fetch stalls any more, but every warp then waits for its memory at the same time and
// the barriers cost more than the fetches did -- measured slower, kept as an experiment.)  have = false: no read for this warp.
// Returns 0 = done, 1 = the read has to take the generic path instead (nothing has been written then).
template <bool TRIM, bool PILE>
AMP_WD int ont_read(const KParams& P, const OntMem& wm, int* cnt, int* ctrl, int wt, int wbase, long long i, bool have, int lane,
                    uint32_t& parity, bool fast_ok) {
    int st = have ? 0 : 2;                            // 0: in progress, 1: declined, 2: nothing (more) to do
    uint32_t c0 = 0, qo0 = 0, so0 = 0, q_lo = 0, s_lo = 0;
    int flag = 0, pos = 0, nc = 0, l_seq = 0;
    int hl = 0, ht = 0, lead = 0, trail = 0, ka = 0, kb = 0, fc = 0, bk = 0, pp = 0, f = 0;
    bool rev = false;
    const uint32_t* ops = wm.ops; const int* q0 = wm.q0; const int* r0 = wm.r0;
    const int minq = P.tp.min_quality;
    // ---- step 1: the read's ops and their prefix sums, the two primer clips ---------------------------------------------------
    if (st == 0) {
        c0 = P.b.cig_off[i]; const uint32_t c1 = P.b.cig_off[i + 1];
        qo0 = P.b.qual_off[i]; const uint32_t qo1 = P.b.qual_off[i + 1];
        uint32_t so1 = 0;
        if (PILE) { so0 = P.b.seq_off[i]; so1 = P.b.seq_off[i + 1]; }
        flag = P.b.flag[i]; pos = P.b.pos[i];
        nc = (int)(c1 - c0); l_seq = (int)(qo1 - qo0);
        rev = (flag & 16) != 0;
        const uint32_t* cig = P.b.cigar + c0;
        if ((flag & 4) || nc == 0) {                                                // AmpliPy.py:902
            if (TRIM) {
                uint32_t* orow = P.o.cigar + (size_t)c0 + 3 * (size_t)i;
                for (int k = lane; k < nc; k += 32) orow[k] = cig[k];
                if (lane == 0) { P.o.pos[i] = pos; P.o.ncig[i] = (uint16_t)nc; P.o.flags[i] = (uint8_t)AMP_F_SKIPPED; }
            }
            st = 2;
        } else if (!fast_ok || nc > AMPO_MAXOPS || l_seq > AMPO_QCAP || l_seq < 1) st = 1;
        if (st == 0) {
            // rows towards shared memory (whole 16-byte pieces), ops + prefix sums meanwhile
            q_lo = qo0 & ~15u; s_lo = so0 & ~15u;
            const uint32_t q_bulk = (qo1 - q_lo + 15u) & ~15u, s_bulk = PILE ? (so1 - s_lo + 15u) & ~15u : 0u;
            if (lane == 0) {
                bulk_expect(wm.bar, q_bulk + s_bulk);
                bulk_copy(wm.qbuf + AMP7_PAD, P.b.qual + q_lo, q_bulk, wm.bar);
                if (s_bulk) bulk_copy(wm.sbuf + AMP7_PAD, P.b.seq + s_lo, s_bulk, wm.bar);
            }
            bool bad = false;
            int qtot = 0, rtot = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int base = 0; base < nc; base += 32) {
                const int k = base + lane;
                const uint32_t w = k < nc ? cig[k] : 0u;
                const uint32_t prev = k > 0 && k < nc ? cig[k - 1] : 0xFu;
                const uint32_t op = c_op(w);
                const int n = c_len(w);
                if (k < nc) {
                    wm.ops[k] = w;
                    if (op > OP_X || op == OP_P || n == 0 || c_op(prev) == op) bad = true;
                }
                const int cq = (k < nc && cons_q(op)) ? n : 0, cr = (k < nc && cons_r(op)) ? n : 0;
                int iq = cq, ir = cr;
                for (int d = 1; d < 32; d <<= 1) {
                    const int tq = w_shfl(iq, lane - d), tr = w_shfl(ir, lane - d);
                    if (lane >= d) { iq += tq; ir += tr; }
                }
                if (k < nc) { wm.q0[k] = qtot + iq - cq; wm.r0[k] = rtot + ir - cr; }
                qtot += w_shfl(iq, 31); rtot += w_shfl(ir, 31);
            }
            if (lane == 0) { wm.q0[nc] = qtot; wm.r0[nc] = rtot; }
            w_sync();
            // leading / trailing clips; the ops in between must be M I D N = X
            kb = nc - 1;
            if (c_op(ops[ka]) == OP_H) { hl = c_len(ops[ka]); ++ka; }
            if (ka < nc && c_op(ops[ka]) == OP_S) { lead = c_len(ops[ka]); ++ka; }
            if (kb >= 0 && c_op(ops[kb]) == OP_H) { ht = c_len(ops[kb]); --kb; }
            if (kb >= 0 && c_op(ops[kb]) == OP_S) { trail = c_len(ops[kb]); --kb; }
            for (int base = ka; base <= kb; base += 32) {
                const int k = base + lane;
                if (k <= kb) { const uint32_t op = c_op(ops[k]); if (op == OP_S || op == OP_H) bad = true; }
            }
            bad = w_ballot(bad) != 0u;
            bool ok = !bad && ka <= kb && qtot == l_seq && cons_qr(c_op(ops[ka < nc ? ka : 0])) && cons_qr(c_op(ops[kb >= 0 ? kb : 0])) &&
                      pos >= 0 && pos + rtot <= P.tp.L;
            bk = ok ? c_len(ops[kb]) : 0; pp = pos;
            if (TRIM && ok) {
                const bool paired = flag & 1;
                const int tlen = P.b.tlen[i];
                const int L1 = P.tp.max_primer_end[pos];                            // 450
                const int R1 = P.tp.min_primer_start[pos + rtot - 1];               // 451
                const int abs_tlen = tlen < 0 ? -tlen : tlen;
                const bool isize = (abs_tlen - P.tp.max_primer_len) > l_seq;        // 452
                const int ka0 = ka, kb0 = kb;
                if (!(paired && isize && rev) && L1 >= 0) {                         // 460
                    const int T = L1 + 1;
                    if (T - pos < 1) ok = false;
                    int ks = -1;
                    if (ok) ks = first_op(ka0, kb0, lane, [&](int k) { return cons_r(c_op(ops[k])) && T <= pos + r0[k] + c_len(ops[k]); });
                    if (ks < 0) ok = false;
                    if (ok) {
                        const int off = T - (pos + r0[ks]);
                        if (cons_q(c_op(ops[ks])) && off < c_len(ops[ks])) { ka = ks; fc = off; }
                        else {
                            ka = first_op(ks + 1, kb0, lane, [&](int k) { return cons_qr(c_op(ops[k])); });
                            fc = 0;
                            if (ka < 0) { ok = false; ka = ka0; }
                        }
                    }
                    if (ok) { lead = q0[ka] + fc; pp = pos + r0[ka] + fc; hl = 0; f |= AMP_F_TRIM_START; }
                }
                if (ok && !(paired && isize && !rev) && R1 >= 0) {                  // 517
                    if (R1 - pp < 1) ok = false;
                    int ks = -1;
                    if (ok) ks = first_op(ka, kb0, lane, [&](int k) { return cons_r(c_op(ops[k])) && R1 <= pos + r0[k] + c_len(ops[k]); });
                    if (ks < 0) ok = false;
                    if (ok) {
                        if (cons_q(c_op(ops[ks]))) { kb = ks; bk = R1 - (pos + r0[ks]); }
                        else {
                            kb = last_op(ka, ks - 1, lane, [&](int k) { return cons_qr(c_op(ops[k])); });
                            if (kb < 0) { ok = false; kb = kb0; } else bk = c_len(ops[kb]);
                        }
                    }
                    if (ok && kb == ka && bk <= fc) ok = false;
                    if (ok) { trail = l_seq - (q0[kb] + bk); ht = 0; f |= AMP_F_TRIM_END; }
                }
            }
            // the staged rows are needed from here on; a declined read leaves the barrier in a known phase
            bulk_wait(wm.bar, parity); parity ^= 1u;
            if (!ok) st = 1;
        }
    }
    AMPO_STEP_SYNC();
    // ---- step 2: sliding-window search, quality clip, write gate, the trimmed read's outputs ---------------------------------------
    const int qrow = (int)(AMP7_PAD + (qo0 - q_lo));                                // the read's first quality byte in qbuf
    bool empty = false;
    auto estart = [&](int k) { return k == ka ? q0[k] + fc : q0[k]; };
    auto eend = [&](int k) { return k == kb ? q0[k] + bk : q0[k] + c_len(ops[k]); };
    auto elen = [&](int k) { return (k == kb ? bk : c_len(ops[k])) - (k == ka ? fc : 0); };   // effective length of op k of the range
    int ref_len1 = 1, rstart = 0;
    if (st == 0) {
        if (TRIM) {
            const int m = l_seq - trail - lead;                                     // aligned query bases (561-563)
            const int del0 = warp_window_del(wm.qbuf, qrow + lead, m, rev, minq, lane);   // 566-587 / 628-649
            if (!rev && del0 != 0) {                                                // 656: from the end
                f |= AMP_F_TRIM_QUAL;
                if (del0 >= m) empty = true;
                else {
                    const int Qc = l_seq - trail - del0;                            // first clipped query base
                    const int k3 = last_op(ka, kb, lane, [&](int k) { return cons_q(c_op(ops[k])) && estart(k) < Qc; });
                    if (Qc < eend(k3)) { kb = k3; bk = Qc - q0[k3]; }
                    else {                                                          // on a boundary: deletions behind k3 survive
                        const int k4 = first_op(k3 + 1, kb, lane, [&](int k) { return cons_q(c_op(ops[k])); });
                        const int kn = k4 - 1;
                        bk = kn == k3 ? Qc - q0[k3] : c_len(ops[kn]);
                        kb = kn;
                    }
                }
                trail += del0;
            } else if (rev && del0 >= 2) {                                          // 591-594: from the start, pos stays (F6)
                f |= AMP_F_TRIM_QUAL;
                if (del0 >= m) empty = true;
                else {
                    const int Qc = lead + del0;                                     // first kept query base
                    const int k3 = first_op(ka, kb, lane, [&](int k) { return cons_q(c_op(ops[k])) && eend(k) > Qc; });
                    if (Qc > estart(k3)) { ka = k3; fc = Qc - q0[k3]; }
                    else {
                        const int k4 = last_op(ka, k3 - 1, lane, [&](int k) { return cons_q(c_op(ops[k])); });
                        ka = k4 + 1; fc = 0;
                    }
                }
                lead += del0;
            }
        }
        int ref_len = 0;
        if (!empty) {
            rstart = r0[ka] + (cons_r(c_op(ops[ka])) ? fc : 0);
            ref_len = r0[kb] + (cons_r(c_op(ops[kb])) ? bk : 0) - rstart;
        }
        ref_len1 = ref_len > 0 ? ref_len : 1;                                       // htslib bam_endpos floor
        if (TRIM) {
            if (ref_len1 >= P.tp.min_length && ((f & (AMP_F_TRIM_START | AMP_F_TRIM_END)) || P.tp.include_no_primer)) f |= AMP_F_KEEP;   // 910
            uint32_t* orow = P.o.cigar + (size_t)c0 + 3 * (size_t)i;
            int n0 = 0;
            if (empty) {
                if (lane == 0) {
                    if (hl > 0) orow[n0++] = c_pack(OP_H, hl);
                    orow[n0++] = c_pack(OP_S, l_seq);
                    if (ht > 0) orow[n0++] = c_pack(OP_H, ht);
                    P.o.pos[i] = pp; P.o.ncig[i] = (uint16_t)n0; P.o.flags[i] = (uint8_t)f;
                }
            } else {
                const int pre = (hl > 0) + (lead > 0), nmid = kb - ka + 1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                for (int k = ka + lane; k <= kb; k += 32) orow[pre + k - ka] = c_pack(c_op(ops[k]), elen(k));
                if (lane == 0) {
                    if (hl > 0) orow[n0++] = c_pack(OP_H, hl);
                    if (lead > 0) orow[n0++] = c_pack(OP_S, lead);
                    n0 += nmid;
                    if (trail > 0) orow[n0++] = c_pack(OP_S, trail);
                    if (ht > 0) orow[n0++] = c_pack(OP_H, ht);
                    P.o.pos[i] = pp; P.o.ncig[i] = (uint16_t)n0; P.o.flags[i] = (uint8_t)f;
                }
            }
        }
    }
    AMPO_STEP_SYNC();
    // ---- steps 3 and 4: update_base_counts ----------------------------------------------------------------------------------------
    const bool pile = PILE && st == 0 && !empty;
    const int nrow = (int)(2u * (AMP7_PAD + so0 - s_lo));                           // nibble index of the read's first base in sbuf
    int anchor = pp + ref_len1 - 1; if (anchor < 0) anchor = 0;                     // max(reference_end - 1, 0)
    unsigned errs = 0;
    // (3) aligned bases (718, 752-753): the aligned query positions are dealt out in equal contiguous spans, a lane walks the ops its
    // span crosses (op lengths are far too uneven for an op per lane: the longest of 32 runs is ~4x the mean)
    if (pile) {
        const int qs = lead, qe = l_seq - trail, span = (qe - qs + 31) >> 5;
        int x = qs + lane * span;
        const int xe = x + span < qe ? x + span : qe;
        if (x < xe) {
            int lo = ka, hi = kb;                                                   // last op with q0 <= x: it consumes query (see q0's ties)
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (q0[mid] <= x) lo = mid; else hi = mid - 1; }
            int k = lo;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            while (x < xe) {
                const uint32_t op = c_op(ops[k]);
                const int kend = eend(k);
                const int n = (xe < kend ? xe : kend) - x;
                if (cons_qr(op)) {
                    const int rb = pp + r0[k] + (x - q0[k]) - rstart;
                    const int w0 = rb - wbase;
                    const uint8_t* qp = wm.qbuf + qrow + x;
                    const uint32_t nb0 = (uint32_t)(nrow + x);
                    if (wbase >= 0 && w0 >= 0 && w0 + n <= wt) {
                        int* tl = cnt + w0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                        for (int j = 0; j < n; ++j) {
                            if (qp[j] < minq) continue;
                            const uint32_t nb = nb0 + (uint32_t)j;
                            atomic_add(&tl[(int)((wm.sbuf[nb >> 1] >> ((~nb & 1u) << 2)) & 15u) * wt + j], 1);
                        }
                    } else {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                        for (int j = 0; j < n; ++j) {
                            if (qp[j] < minq) continue;
                            const uint32_t nb = nb0 + (uint32_t)j;
                            count_global7(P, cnt, wbase, (int)((wm.sbuf[nb >> 1] >> ((~nb & 1u) << 2)) & 15u), rb + j, errs);
                        }
                    }
                }
                x += n;
                if (x >= kend) { ++k; while (k <= kb && !cons_q(c_op(ops[k]))) ++k; }
            }
        }
    }
    AMPO_STEP_SYNC();
    // (4) deletions (714-715) and insertions (730-748), an op per lane
    if (pile) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int k = ka + lane; k <= kb; k += 32) {
            const uint32_t op = c_op(ops[k]);
            if (cons_qr(op)) continue;
            const int n = elen(k);
            const int rb = pp + r0[k] - rstart;                                     // (a D / N / I op is never cut)
            if (op == OP_D || op == OP_N) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                for (int j = 0; j < n; ++j) count_global7(P, cnt, wbase, AMP7_DEL_CODE, rb + j, errs);
            } else {
                const int qb = estart(k), qend = qb + n;
                int qi = -1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                for (int j = qb; j < qend; ++j) {
                    const bool pass = wm.qbuf[qrow + j] >= minq;
                    if (qi < 0) { if (pass) qi = j; }
                    else if (!pass) { ont_defer(P, wm, so0, anchor, qi - 1, j - (qi - 1)); qi = -1; }      // exit (D)
                }
                if (qi >= 0) {
                    int ipos = anchor, ib = qi - 1, ie = qend;                      // exit (C): the trailing clip follows
                    if (k < kb) {
                        if (cons_qr(c_op(ops[k + 1]))) {                            // exit (A) / (A')
                            if (rb == 0) { ipos = 0; ib = qi; ie = qend + 1 < l_seq ? qend + 1 : l_seq; }
                            else ipos = rb - 1;
                        } else {                                                    // exit (B): the key runs to the end of the read
                            if (rb == 0) { errs |= AMP_E_INS_END; ie = -1; }
                            else { ipos = rb - 1; ie = l_seq; }
                        }
                    } else if (trail <= 0) { errs |= AMP_E_INS_END; ie = -1; }      // IndexError at 734
                    if (ie >= 0) ont_defer(P, wm, so0, ipos, ib, ie - ib);
                }
            }
        }
    }
    if (errs) atomic_or(P.err, errs);
    w_sync();   // every lane is done with the warp's arrays before the next read overwrites them
    return st == 1 ? 1 : 0;
}

// ---- the kernel: one CTA per SM, a contiguous chunk of the (coordinate-sorted) reads and one count tile per CTA; the warps take
// reads one at a time from a shared counter.  P.tiles_per_cta = reads per CTA, P.ntiles = reads.
template <bool TRIM, bool PILE, int WT>
AMP_WD void cta_trim_pileup_ont(const KParams& P, unsigned char* smem_base, int gwarps) {
    const int wt = WT ? WT : P.wt;
    const int tid = c_tid(), nthreads = c_nthreads(), block = c_block();
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    int* cnt = (int*)smem_base;
    int* ctrl = (int*)(smem_base + tile_bytes_v7(wt));
    const OntMem wm = carve_ont(smem_base, wt, warp);
    const long long r_lo = (long long)block * P.tiles_per_cta;
    long long r_hi = r_lo + P.tiles_per_cta; if (r_hi > P.ntiles) r_hi = P.ntiles;
    const int n_reads = r_hi > r_lo ? (int)(r_hi - r_lo) : 0;
    const long long n_end = P.b.first + P.b.n;
    uint32_t* glist = P.glist + (size_t)block * P.gcap;

    if (PILE) {
        for (int i = tid; i < AMP7_SINK_ROW * wt; i += nthreads) cnt[i] = 0;
        for (int i = tid; i < wt; i += nthreads) cnt[del_row_off(wt) + i] = 0;
    }
    if (tid == 0) { ctrl[C7_NEXT] = 0; ctrl[C7_NGEN] = 0; ctrl[C7_GNEXT] = 0; ctrl[C7_FASTDONE] = 0; ctrl[C7_NEV] = 0; ctrl[C7_EVDONE] = 0; }
    for (int k = tid; k < AMP7_GLCAP; k += nthreads) ctrl[C7_GL + k] = 0;
    for (int k = tid; k < AMP7_EVCAP; k += nthreads) ctrl[C7_EV + 3 * k + 2] = 0;
    if (lane == 0) mbar_init(wm.bar);
    const int wb = chunk_window_base(P, ctrl, P.b.first + r_lo, n_end, tid, n_reads > 0);
    const int wbase = PILE ? wb : -1;
    const int minq = P.tp.min_quality;
    const bool fast_ok = minq >= 0 && minq <= 127 && (!TRIM || P.tp.window == 4);
    uint32_t parity = 0;
    if (lane == 0) *wm.nev = 0;
    w_sync();
#ifdef AMPO_LOCKSTEP
    for (int r = warp; r - warp < n_reads; r += nwarps) {       // rounds of nwarps reads
#else
    for (;;) {
        int r = 0;
        if (lane == 0) r = atomic_add(&ctrl[C7_NEXT], 1);
        r = w_shfl(r, 0);
        if (r >= n_reads) break;
#endif
        const bool have = r < n_reads;
        const long long i = P.b.first + r_lo + (have ? r : 0);
        const int rc = ont_read<TRIM, PILE>(P, wm, cnt, ctrl, wt, wbase, i, have, lane, parity, fast_ok);
#if !defined(__CUDA_ARCH__)
        if (lane == 0 && have) ++g_v7_stats[rc ? 1 : 0];        // emulation only: reads finished by the warp / sent to the generic path
#endif
        if (rc) {
            if (lane == 0) {
                const int idx = atomic_add(&ctrl[C7_NGEN], 1);
                const uint32_t v = (uint32_t)(i - P.b.first) + 1u;
                if (idx < AMP7_GLCAP) st_vol(&ctrl[C7_GL + idx], (int)v); else st_cg_u32(&glist[idx - AMP7_GLCAP], v);
            }
            w_sync();
        }
        if (PILE && *wm.nev >= 32) ont_drain(P, wm, lane);       // a full warp's worth of parked alleles
    }
    if (PILE) ont_drain(P, wm, lane);
    fence_block();
    if (lane == 0) atomic_add(&ctrl[C7_FASTDONE], 1);
    // ---- the listed reads: the loop-for-loop generic path (amp_warp.cuh), AMP7_GN reads per round, by the last warps
    if (warp >= nwarps - gwarps) {
        const WarpMem7 gm = carve_ont_generic(smem_base, wt, nwarps, warp - (nwarps - gwarps));
        if (lane == 0) mbar_init(gm.bar);
        uint32_t gparity = 0;
        for (;;) {
            int at = -1, n = 0;
            if (lane == 0) {
                for (;;) {
                    const bool done = ld_vol(&ctrl[C7_FASTDONE]) >= nwarps;
                    const int reserved = ld_vol(&ctrl[C7_NGEN]), claimed = ld_vol(&ctrl[C7_GNEXT]);
                    int avail = reserved - claimed;
                    if (claimed < AMP7_GLCAP) { if (avail > AMP7_GLCAP - claimed) avail = AMP7_GLCAP - claimed; }
                    else if (!done) avail = 0;
                    if (avail >= AMP7_GN || (done && avail > 0)) {
                        n = avail < AMP7_GN ? avail : AMP7_GN;
                        if (atomic_cas(&ctrl[C7_GNEXT], claimed, claimed + n) == claimed) { at = claimed; break; }
                        continue;
                    }
                    if (done) break;
                    c_yield();
                }
            }
            at = w_shfl(at, 0); n = w_shfl(n, 0);
            if (at < 0) break;
            fence_block();
            if (lane < n) {
                const int idx = at + lane;
                uint32_t v;
                if (idx < AMP7_GLCAP) { while ((v = (uint32_t)ld_vol(&ctrl[C7_GL + idx])) == 0u) c_yield(); }
                else v = ld_cg_u32(&glist[idx - AMP7_GLCAP]);
                gm.queue[lane] = v - 1u;
            }
            w_sync();
            warp_generic_phase<WT>(P, gm, cnt, wt, wbase, n, n, lane, TRIM, PILE, gparity, nullptr);
        }
    }
    if (PILE) ins_drain(P, ctrl, lane);
    c_sync();
    if (PILE) ins_drain(P, ctrl, lane);
    if (PILE && wbase >= 0) flush_tile7(P, cnt, wbase, tid, nthreads);
}

}  // namespace amp
