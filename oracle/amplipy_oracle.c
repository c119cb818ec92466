/*
 * amplipy_oracle.c -- CPU restatement of AmpliPy's trim -> pileup -> call path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path and the
 * "port" CPU baseline of bench.py.  It must never be imported, linked or called by the
 * product package (amplipy_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * Parity pin: every function below is checked against the UNMODIFIED reference
 * (/root/reference/AmpliPy.py, executed through oracle/pysam_shim) by
 * tests/golden/make_golden.py (fixtures committed under tests/golden/) and by
 * tests/test_oracle_vs_reference.py (runs where the reference is mounted).
 * The reference ships no tests of its own; pysam==0.17.0 semantics
 * (requirements.txt:1) are restated from its documentation -- see oracle/pysam_shim/pysam.py.
 *
 * Each function cites the reference lines it restates.  The control flow deliberately
 * mirrors the reference loop-for-loop (including its quirks) instead of using closed forms;
 * the closed forms live in the CUDA kernels and are validated against this file.
 *
 * Data layout = amplipy_b200/batch.py (struct-of-arrays; BAM packed CIGAR words; 4-bit seq).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OP_M 0
#define OP_I 1
#define OP_D 2
#define OP_N 3
#define OP_S 4
#define OP_H 5
#define OP_P 6
#define OP_EQ 7
#define OP_X 8

/* AmpliPy.py:43-44 */
static const int CONSUME_QUERY[16] = {1, 1, 0, 0, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const int CONSUME_REF[16]   = {1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0};

/* flag bits written to out_flags (shared with include/amplipy_b200.h) */
#define F_TRIM_START 1
#define F_TRIM_END   2
#define F_TRIM_QUAL  4
#define F_KEEP       8
#define F_SKIPPED    16
#define F_ERROR      32

typedef struct { int op; int n; } cig_t;

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* fix the number of OpenMP threads whatever OMP_NUM_THREADS says (bench.py: the same count at every --gpus N) */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- AmpliPy.py:174-209 find_overlapping_primers (deque emulated with head/tail indices) ---- */
int oracle_find_overlapping_primers(int L, int P, const int32_t* starts, const int32_t* ends, int offset,
                                    int32_t* min_start, int32_t* max_end) {
    int head = 0, i = 0; /* deque = primers[head..i) */
    for (int p = 0; p < L; ++p) {
        while (head != i && p >= ends[head] + offset) head++;
        while (i < P && p >= starts[i] - offset) i++;
        if (head != i) {
            int mn = starts[head], mx = ends[head];
            for (int k = head; k < i; ++k) { if (starts[k] < mn) mn = starts[k]; if (ends[k] > mx) mx = ends[k]; }
            min_start[p] = mn; max_end[p] = mx;
        } else { min_start[p] = -1; max_end[p] = -1; }
    }
    return 0;
}

/* ---- AmpliPy.py:363-386 ---- */
static int get_pos_on_ref(const cig_t* c, int nc, int query_pos, int ref_start) {
    int cur_pos = 0, ref_pos = ref_start;
    for (int k = 0; k < nc; ++k) {
        int cig = c[k].op, n = c[k].n;
        if (CONSUME_QUERY[cig]) {
            if (query_pos <= cur_pos + n) {
                if (CONSUME_REF[cig]) ref_pos += (query_pos - cur_pos);
                return ref_pos;
            }
            cur_pos += n;
        }
        if (CONSUME_REF[cig]) ref_pos += n;
    }
    return ref_pos;
}

/* ---- AmpliPy.py:389-412 ---- */
static int get_pos_on_query(const cig_t* c, int nc, int ref_pos, int ref_start) {
    int query_pos = 0, cur_pos = ref_start;
    for (int k = 0; k < nc; ++k) {
        int cig = c[k].op, n = c[k].n;
        if (CONSUME_REF[cig]) {
            if (ref_pos <= cur_pos + n) {
                if (CONSUME_QUERY[cig]) query_pos += (ref_pos - cur_pos);
                return query_pos;
            }
            cur_pos += n;
        }
        if (CONSUME_QUERY[cig]) query_pos += n;
    }
    return query_pos;
}

/* ---- AmpliPy.py:415-423 (in place; returns new length) ---- */
static int fix_cigar(cig_t* c, int nc) {
    int out = 0;
    for (int i = 0; i < nc; ++i) {
        if (i < nc - 1 && c[i].op == c[i + 1].op) { c[i + 1].n += c[i].n; continue; }
        c[out++] = c[i];
    }
    return out;
}

static void reverse_cigar(cig_t* c, int nc) {
    for (int i = 0, j = nc - 1; i < j; ++i, --j) { cig_t t = c[i]; c[i] = c[j]; c[j] = t; }
}

/* pysam semantics restated in oracle/pysam_shim/pysam.py */
static int ref_len_of(const cig_t* c, int nc) {
    int r = 0;
    for (int k = 0; k < nc; ++k) if (CONSUME_REF[c[k].op]) r += c[k].n;
    return r == 0 ? 1 : r; /* htslib bam_endpos floor */
}
static int query_alignment_start(const cig_t* c, int nc) {
    int s = 0;
    for (int k = 0; k < nc; ++k) {
        if (c[k].op == OP_H) continue;
        else if (c[k].op == OP_S) s += c[k].n;
        else break;
    }
    return s;
}
static int query_alignment_end(const cig_t* c, int nc, int l_seq) {
    int e = l_seq;
    for (int k = nc - 1; k >= 1; --k) {
        if (c[k].op == OP_H) continue;
        else if (c[k].op == OP_S) e -= c[k].n;
        else break;
    }
    return e;
}

/*
 * ---- AmpliPy.py:426-687 trim_read.  c/nc = working CIGAR (capacity nc+3), a = scratch of the same
 * capacity.  Returns flag bits; updates *pos and *pnc. ----
 */
static int trim_read(cig_t* c, int* pnc, cig_t* a, int32_t* pos, int flag, int tlen, int l_seq,
                     const uint8_t* qual_all, int L, const int32_t* min_primer_start, const int32_t* max_primer_end,
                     int max_primer_len, int min_quality, int sliding_window_width) {
    int nc = *pnc;
    int ref_start = *pos;
    int is_paired = flag & 1, is_reverse = (flag & 16) != 0;
    int ref_end = ref_start + ref_len_of(c, nc);
    if (ref_start < 0 || ref_start >= L || ref_end - 1 >= L) return F_ERROR; /* IndexError in the reference (450-451) */
    int left_max_primer_end = max_primer_end[ref_start];              /* 450 */
    int right_min_primer_start = min_primer_start[ref_end - 1];       /* 451 */
    int abs_tlen = tlen < 0 ? -tlen : tlen;
    int isize_flag = (abs_tlen - max_primer_len) > l_seq;             /* 452 */
    int out = 0;

    /* 460-514: primer at the start */
    if (!(is_paired && isize_flag && is_reverse) && left_max_primer_end >= 0) {
        out |= F_TRIM_START;
        int del_len = get_pos_on_query(c, nc, left_max_primer_end + 1, ref_start);
        int na = 0, ref_add = 0, pos_start = 0, start_pos = 0;
        for (int k = 0; k < nc; ++k) {
            int cig = c[k].op, n = c[k].n;
            if (del_len == 0 && pos_start) { a[na++] = c[k]; continue; }
            if (del_len == 0 && CONSUME_QUERY[cig] && CONSUME_REF[cig]) { pos_start = 1; a[na++] = c[k]; continue; }
            ref_add = 0;
            if (CONSUME_QUERY[cig]) {
                if (del_len >= n) { a[na].op = OP_S; a[na].n = n; na++; }
                else if (0 < del_len && del_len < n) { a[na].op = OP_S; a[na].n = del_len; na++; }
                else { a[na].op = OP_S; a[na].n = n; na++; continue; }
                ref_add = del_len < n ? del_len : n;
                int tmp = n;
                n = (n - del_len) > 0 ? (n - del_len) : 0;
                del_len = (del_len - tmp) > 0 ? (del_len - tmp) : 0;
                if (n > 0) { a[na].op = cig; a[na].n = n; na++; }
                if (del_len == 0 && CONSUME_QUERY[a[na - 1].op] && CONSUME_REF[a[na - 1].op]) pos_start = 1;
            } else if (CONSUME_REF[cig]) {
                ref_add += n;
            }
            if (CONSUME_REF[cig]) start_pos += ref_add;
        }
        nc = fix_cigar(a, na);
        memcpy(c, a, sizeof(cig_t) * (size_t)nc);
        ref_start += start_pos;
    }

    /* 516-558: primer at the end */
    if (!(is_paired && isize_flag && !is_reverse) && right_min_primer_start >= 0) {
        out |= F_TRIM_END;
        int del_len = l_seq - get_pos_on_query(c, nc, right_min_primer_start, ref_start);
        int na = 0, pos_start = 0;
        for (int k = nc - 1; k >= 0; --k) {
            int cig = c[k].op, n = c[k].n;
            if (del_len == 0 && pos_start) { a[na++] = c[k]; continue; }
            if (del_len == 0 && CONSUME_QUERY[cig] && CONSUME_REF[cig]) { pos_start = 1; a[na++] = c[k]; continue; }
            if (CONSUME_QUERY[cig]) {
                if (del_len >= n) { a[na].op = OP_S; a[na].n = n; na++; }
                else if (0 < del_len && del_len < n) { a[na].op = OP_S; a[na].n = del_len; na++; }
                else { a[na].op = OP_S; a[na].n = n; na++; continue; }
                int tmp = n;
                n = (n - del_len) > 0 ? (n - del_len) : 0;
                del_len = (del_len - tmp) > 0 ? (del_len - tmp) : 0;
                if (n > 0) { a[na].op = cig; a[na].n = n; na++; }
                if (del_len == 0 && CONSUME_QUERY[a[na - 1].op] && CONSUME_REF[a[na - 1].op]) pos_start = 1;
            }
        }
        reverse_cigar(a, na);
        nc = fix_cigar(a, na);
        memcpy(c, a, sizeof(cig_t) * (size_t)nc);
    }

    /* 560-563: quality window set-up on query_alignment_qualities */
    int qas = query_alignment_start(c, nc);
    int qae = query_alignment_end(c, nc, l_seq);
    const uint8_t* qual = qual_all + qas;
    int true_start = 0, true_end = qae - qas;
    if (true_end < 0) true_end = 0;
    long total = 0;
    int window = sliding_window_width < true_end ? sliding_window_width : true_end;

    if (is_reverse) {
        /* 566-625 */
        int i = true_end;
        for (int offset = 1; offset < window; ++offset) total += qual[i - offset];
        while (i > true_start) {
            if (true_start + window > i) window -= 1;
            else total += qual[i - window];
            if ((double)total / (double)window < (double)min_quality) break;
            total -= qual[i - 1]; i -= 1;
        }
        int del_len = i;
        int start_pos = get_pos_on_ref(c, nc, del_len + qas - 1, ref_start);
        if (start_pos > ref_start) {
            out |= F_TRIM_QUAL;
            int na = 0;
            for (int k = 0; k < nc; ++k) {
                int cig = c[k].op, n = c[k].n;
                if (del_len == 0) { a[na++] = c[k]; continue; }
                if (cig == OP_S || cig == OP_H) { a[na++] = c[k]; continue; }
                if (CONSUME_QUERY[cig]) {
                    if (del_len >= n) { a[na].op = OP_S; a[na].n = n; na++; }
                    else { a[na].op = OP_S; a[na].n = del_len; na++; }
                    int tmp = n;
                    n = (n - del_len) > 0 ? (n - del_len) : 0;
                    del_len = (del_len - tmp) > 0 ? (del_len - tmp) : 0;
                    if (n > 0) { a[na].op = cig; a[na].n = n; na++; }
                }
            }
            nc = fix_cigar(a, na);
            memcpy(c, a, sizeof(cig_t) * (size_t)nc);
            /* NOTE: reference_start is NOT advanced here (AmpliPy.py:589-625) */
        }
    } else {
        /* 628-686 */
        int i = true_start;
        for (int offset = 0; offset < window - 1; ++offset) total += qual[i + offset];
        while (i < true_end) {
            if ((true_end - window) < i) window -= 1;
            else total += qual[i + window - 1];
            if ((double)total / (double)window < (double)min_quality) break;
            total -= qual[i]; i += 1;
        }
        int del_len = true_end - i;
        if (del_len != 0) {
            out |= F_TRIM_QUAL;
            int na = 0;
            for (int k = nc - 1; k >= 0; --k) {
                int cig = c[k].op, n = c[k].n;
                if (del_len == 0) { a[na++] = c[k]; continue; }
                if (cig == OP_S || cig == OP_H) { a[na++] = c[k]; continue; }
                if (CONSUME_QUERY[cig]) {
                    if (del_len >= n) { a[na].op = OP_S; a[na].n = n; na++; }
                    else { a[na].op = OP_S; a[na].n = del_len; na++; }
                    int tmp = n;
                    n = (n - del_len) > 0 ? (n - del_len) : 0;
                    del_len = (del_len - tmp) > 0 ? (del_len - tmp) : 0;
                    if (n > 0) { a[na].op = cig; a[na].n = n; na++; }
                }
            }
            reverse_cigar(a, na);
            nc = fix_cigar(a, na);
            memcpy(c, a, sizeof(cig_t) * (size_t)nc);
        }
    }
    *pnc = nc;
    *pos = ref_start;
    return out;
}

/*
 * Batch driver for trimming: AmpliPy.py:896-911 (skip rule 902, write gate 910).
 * out_cigar row i starts at cig_off[i] + 3*i and has capacity n_cigar[i] + 3.
 */
int oracle_trim_batch(int64_t N, const int32_t* pos, const uint16_t* flag, const int32_t* tlen,
                      const uint32_t* cig_off, const uint32_t* cigar, const uint32_t* qual_off, const uint8_t* qual,
                      int L, const int32_t* min_primer_start, const int32_t* max_primer_end, int max_primer_len,
                      int min_quality, int sliding_window_width, int min_length, int include_no_primer,
                      int32_t* out_pos, int32_t* out_ncig, uint32_t* out_cigar, uint8_t* out_flags) {
#pragma omp parallel
    {
        cig_t* c = NULL; cig_t* a = NULL; int cap = 0;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            int nc = (int)(cig_off[i + 1] - cig_off[i]);
            uint32_t* orow = out_cigar + (size_t)cig_off[i] + 3 * (size_t)i;
            out_pos[i] = pos[i];
            if ((flag[i] & 4) || nc == 0) { out_ncig[i] = nc; out_flags[i] = F_SKIPPED;
                for (int k = 0; k < nc; ++k) orow[k] = cigar[cig_off[i] + k];
                continue; }
            if (nc + 3 > cap) { cap = 2 * (nc + 3); c = (cig_t*)realloc(c, sizeof(cig_t) * cap); a = (cig_t*)realloc(a, sizeof(cig_t) * cap); }
            for (int k = 0; k < nc; ++k) { uint32_t w = cigar[cig_off[i] + k]; c[k].op = (int)(w & 15); c[k].n = (int)(w >> 4); }
            int32_t p = pos[i];
            int l_seq = (int)(qual_off[i + 1] - qual_off[i]);
            int f = trim_read(c, &nc, a, &p, flag[i], tlen[i], l_seq, qual + qual_off[i], L, min_primer_start,
                              max_primer_end, max_primer_len, min_quality, sliding_window_width);
            if (f & F_ERROR) { out_ncig[i] = 0; out_flags[i] = F_ERROR; continue; }
            int reference_length = ref_len_of(c, nc);
            if (reference_length >= min_length && ((f & (F_TRIM_START | F_TRIM_END)) || include_no_primer)) f |= F_KEEP; /* 910 */
            out_pos[i] = p; out_ncig[i] = nc; out_flags[i] = (uint8_t)f;
            for (int k = 0; k < nc; ++k) orow[k] = ((uint32_t)c[k].n << 4) | (uint32_t)c[k].op;
        }
        free(c); free(a);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Pileup: AmpliPy.py:690-753.  Fixed symbols in channel order A C G T N '-'.  Insertion alleles are
 * accumulated in a (pos, string) -> count hash table (the reference's per-position dict keys).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t cap, used;
    int32_t* pos; int64_t* off; int32_t* len; int64_t* count;   /* slot arrays; len < 0 = empty */
    char* arena; int64_t arena_cap, arena_used;
} ins_table_t;

static uint64_t hash_key(int32_t pos, const char* s, int len) {
    uint64_t h = 1469598103934665603ULL ^ (uint64_t)(uint32_t)pos;
    h *= 1099511628211ULL;
    for (int i = 0; i < len; ++i) { h ^= (uint8_t)s[i]; h *= 1099511628211ULL; }
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
    return h;
}
static void table_init(ins_table_t* t, int64_t cap) {
    t->cap = cap; t->used = 0;
    t->pos = (int32_t*)malloc(sizeof(int32_t) * cap); t->off = (int64_t*)malloc(sizeof(int64_t) * cap);
    t->len = (int32_t*)malloc(sizeof(int32_t) * cap); t->count = (int64_t*)malloc(sizeof(int64_t) * cap);
    for (int64_t i = 0; i < cap; ++i) t->len[i] = -1;
    t->arena_cap = 1 << 16; t->arena_used = 0; t->arena = (char*)malloc(t->arena_cap);
}
static void table_free(ins_table_t* t) { free(t->pos); free(t->off); free(t->len); free(t->count); free(t->arena); }
static void table_add(ins_table_t* t, int32_t pos, const char* s, int len, int64_t cnt);
static void table_grow(ins_table_t* t) {
    ins_table_t n; table_init(&n, t->cap * 2);
    for (int64_t i = 0; i < t->cap; ++i) if (t->len[i] >= 0) table_add(&n, t->pos[i], t->arena + t->off[i], t->len[i], t->count[i]);
    table_free(t); *t = n;
}
static void table_add(ins_table_t* t, int32_t pos, const char* s, int len, int64_t cnt) {
    if (t->used * 2 >= t->cap) table_grow(t);
    uint64_t h = hash_key(pos, s, len);
    int64_t i = (int64_t)(h & (uint64_t)(t->cap - 1));
    for (;;) {
        if (t->len[i] < 0) {
            while (t->arena_used + len > t->arena_cap) { t->arena_cap *= 2; t->arena = (char*)realloc(t->arena, t->arena_cap); }
            memcpy(t->arena + t->arena_used, s, (size_t)len);
            t->pos[i] = pos; t->off[i] = t->arena_used; t->len[i] = len; t->count[i] = cnt; t->arena_used += len; t->used++;
            return;
        }
        if (t->pos[i] == pos && t->len[i] == len && memcmp(t->arena + t->off[i], s, (size_t)len) == 0) { t->count[i] += cnt; return; }
        i = (i + 1) & (t->cap - 1);
    }
}

static const char NIB2CHAR[17] = "=ACMGRSVTWYHKDBN";
static inline int sym_channel(char ch) {
    switch (ch) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4; default: return -1; }
}

/* one read: literal pair-list walk (706-753).  Returns 0 ok, F_ERROR for the reference's crash cases. */
static int pileup_read(int64_t* counts, int L, ins_table_t* tab, int32_t ref_start, const cig_t* c, int nc,
                       const char* query_seq, const uint8_t* query_qual, int l_seq, int min_quality,
                       int32_t* pq, int32_t* pr) {
    int query_start = query_alignment_start(c, nc);         /* 700 */
    int query_end = query_alignment_end(c, nc, l_seq);      /* 701 */
    int ref_end = ref_start + ref_len_of(c, nc);            /* 705 */
    /* 706 get_aligned_pairs: -1 == None */
    int np = 0, q = 0, r = ref_start;
    for (int k = 0; k < nc; ++k) {
        int op = c[k].op, n = c[k].n;
        if (op == OP_M || op == OP_EQ || op == OP_X) { for (int j = 0; j < n; ++j) { pq[np] = q + j; pr[np] = r + j; np++; } q += n; r += n; }
        else if (op == OP_I || op == OP_S || op == OP_P) { for (int j = 0; j < n; ++j) { pq[np] = q + j; pr[np] = -1; np++; } q += n; }
        else if (op == OP_D || op == OP_N) { for (int j = 0; j < n; ++j) { pq[np] = -1; pr[np] = r + j; np++; } r += n; }
    }
    if (q > l_seq) return F_ERROR;
    if (r > L) return F_ERROR;
    int i = 0;
    while (i < np) {
        int q_pos = pq[i], r_pos = pr[i]; i++;
        if (q_pos < 0) { counts[5 * (int64_t)L + r_pos] += 1; }                  /* 714-715 */
        else if (query_qual[q_pos] < min_quality) continue;                     /* 718 */
        else if (q_pos < query_start) continue;                                 /* 722 */
        else if (q_pos >= query_end) break;                                     /* 726 */
        else if (r_pos < 0) {                                                   /* 730-748 */
            int q0 = q_pos;
            while (r_pos < 0 && q_pos < query_end && query_qual[q_pos] >= min_quality) {
                if (i >= np) return F_ERROR;                                    /* IndexError (734) */
                q_pos = pq[i]; r_pos = pr[i]; i++;
                if (q_pos < 0) break;  /* (None, r): python's `r_pos is None` test fails first */
            }
            int s_begin, s_end;  /* python slice of query_seq -> [s_begin, s_end) */
            if (r_pos == 0) {
                if (q_pos < 0) return F_ERROR;                                  /* None + 1 -> TypeError */
                s_begin = q0; s_end = q_pos + 1;                                /* 736 */
            } else {
                s_begin = q0 - 1; if (s_begin < 0) s_begin += l_seq;            /* negative index wraps (738) */
                if (s_begin < 0) s_begin = 0;
                s_end = (q_pos < 0) ? l_seq : q_pos;
            }
            if (s_end > l_seq) s_end = l_seq;
            int slen = s_end - s_begin; if (slen < 0) slen = 0;
            int ref_insertion_pos;
            if (r_pos < 0) ref_insertion_pos = ref_end;                         /* 739-740 */
            else { ref_insertion_pos = r_pos; i -= 1; }                         /* 742-743 */
            ref_insertion_pos = ref_insertion_pos - 1 > 0 ? ref_insertion_pos - 1 : 0;   /* 744 */
            if (ref_insertion_pos >= L) return F_ERROR;
            /* 745-748: the key goes into the same per-position dict as the fixed symbols, so a
             * one-character key (only reachable via q0 == 0 followed by a deletion: seq[-1:]) lands
             * in that base's own counter. */
            int one = (slen == 1) ? sym_channel(query_seq[s_begin]) : -1;
            if (one >= 0) counts[one * (int64_t)L + ref_insertion_pos] += 1;
            else table_add(tab, ref_insertion_pos, query_seq + s_begin, slen, 1);
        } else {
            if (query_qual[q_pos] >= min_quality) {                             /* 752-753 */
                int ch = sym_channel(query_seq[q_pos]);
                if (ch < 0) return F_ERROR;                                     /* KeyError */
                counts[ch * (int64_t)L + r_pos] += 1;
            }
        }
    }
    return 0;
}

typedef struct { ins_table_t tab; } oracle_pileup_t;

/*
 * Pile up N reads into counts[6][L] (int64, channel-major) and an insertion table handle.
 * ncig/cigar_rows: if row_stride3 != 0 the CIGAR of read i starts at cig_off[i] + 3*i with
 * length ncig[i] (= the trim output layout); otherwise at cig_off[i] with cig_off[i+1]-cig_off[i] ops.
 * skip_flags (may be NULL): reads with F_SKIPPED/F_ERROR are ignored (AmpliPy.py:902).
 * Returns a handle to be read with oracle_ins_* and freed with oracle_pileup_free.
 */
void* oracle_pileup_batch(int64_t N, const int32_t* pos, const uint16_t* flag, const uint32_t* cig_off, const uint32_t* cigar,
                          const int32_t* ncig, int row_stride3, const uint8_t* skip_flags,
                          const uint32_t* seq_off, const uint8_t* seq, const uint32_t* qual_off, const uint8_t* qual,
                          int L, int min_quality, int64_t* counts, int64_t* n_errors) {
    oracle_pileup_t* H = (oracle_pileup_t*)malloc(sizeof(oracle_pileup_t));
    table_init(&H->tab, 1 << 12);
    int64_t errors = 0;
#pragma omp parallel reduction(+ : errors)
    {
        int nthreads = 1, tid = 0;
#ifdef _OPENMP
        nthreads = omp_get_num_threads(); tid = omp_get_thread_num();
#endif
        int64_t* my_counts = counts;
        ins_table_t my_tab; ins_table_t* tab = &H->tab;
        if (nthreads > 1) { my_counts = (int64_t*)calloc((size_t)6 * L, sizeof(int64_t)); table_init(&my_tab, 1 << 12); tab = &my_tab; }
        cig_t* c = NULL; int ccap = 0; char* s = NULL; int32_t* pq = NULL; int32_t* pr = NULL; int scap = 0, pcap = 0;
        int64_t lo = N * tid / nthreads, hi = N * (tid + 1) / nthreads;
        for (int64_t i = lo; i < hi; ++i) {
            if (skip_flags && (skip_flags[i] & (F_SKIPPED | F_ERROR))) continue;
            const uint32_t* row; int nc;
            if (row_stride3) { row = cigar + (size_t)cig_off[i] + 3 * (size_t)i; nc = ncig[i]; }
            else { row = cigar + cig_off[i]; nc = (int)(cig_off[i + 1] - cig_off[i]); }
            if ((flag[i] & 4) || nc == 0) continue;                              /* 902 */
            int l_seq = (int)(qual_off[i + 1] - qual_off[i]);
            if (nc > ccap) { ccap = 2 * nc; c = (cig_t*)realloc(c, sizeof(cig_t) * ccap); }
            int npairs = 0;
            for (int k = 0; k < nc; ++k) { c[k].op = (int)(row[k] & 15); c[k].n = (int)(row[k] >> 4); if (c[k].op != OP_H) npairs += c[k].n; }
            if (l_seq + 1 > scap) { scap = 2 * (l_seq + 1); s = (char*)realloc(s, scap); }
            if (npairs + 1 > pcap) { pcap = 2 * (npairs + 1); pq = (int32_t*)realloc(pq, sizeof(int32_t) * pcap); pr = (int32_t*)realloc(pr, sizeof(int32_t) * pcap); }
            const uint8_t* sp = seq + seq_off[i];
            for (int j = 0; j < l_seq; ++j) s[j] = NIB2CHAR[(sp[j >> 1] >> ((~j & 1) << 2)) & 15];   /* 702 */
            if (pileup_read(my_counts, L, tab, pos[i], c, nc, s, qual + qual_off[i], l_seq, min_quality, pq, pr)) errors++;
        }
        free(c); free(s); free(pq); free(pr);
        if (nthreads > 1) {
#pragma omp critical
            {
                for (int64_t k = 0; k < (int64_t)6 * L; ++k) counts[k] += my_counts[k];
                for (int64_t k = 0; k < my_tab.cap; ++k) if (my_tab.len[k] >= 0)
                    table_add(&H->tab, my_tab.pos[k], my_tab.arena + my_tab.off[k], my_tab.len[k], my_tab.count[k]);
            }
            free(my_counts); table_free(&my_tab);
        }
    }
    if (n_errors) *n_errors = errors;
    return H;
}

int64_t oracle_ins_count(void* h) { return ((oracle_pileup_t*)h)->tab.used; }
int64_t oracle_ins_chars(void* h) { return ((oracle_pileup_t*)h)->tab.arena_used; }
/* export unique insertion alleles: pos[K], count[K], str_off[K+1], chars */
void oracle_ins_export(void* h, int32_t* pos, int64_t* count, int64_t* str_off, char* chars) {
    ins_table_t* t = &((oracle_pileup_t*)h)->tab;
    int64_t k = 0, o = 0;
    for (int64_t i = 0; i < t->cap; ++i) if (t->len[i] >= 0) {
        pos[k] = t->pos[i]; count[k] = t->count[i]; str_off[k] = o;
        memcpy(chars + o, t->arena + t->off[i], (size_t)t->len[i]); o += t->len[i]; k++;
    }
    str_off[k] = o;
}
void oracle_pileup_free(void* h) { table_free(&((oracle_pileup_t*)h)->tab); free(h); }

/* ------------------------------------------------------------------------------------------------
 * Calling: AmpliPy.py:756-771 (alleles_from_counts), 919-929 (consensus), 932-951 (variants).
 * Insertion alleles arrive sorted by position: ins_pos[K] ascending, with strings/ counts.
 * Output (flattened):
 *   depth[L]; al_off[L+1]; per allele (sorted as the reference sorts): al_count, al_freq, al_sym
 *   where al_sym = 0..5 for A,C,G,T,N,'-' and 6+k for insertion allele k; al_is_alt flags;
 *   cons_sym[L] = allele id of the consensus symbol or -1 for unknown_symbol;
 *   var_emit[L], var_gt_ref[L], var_ref_count[L], var_ref_freq[L].
 * ------------------------------------------------------------------------------------------------ */
typedef struct { int64_t count; double freq; const char* s; int len; int id; } allele_t;
static int cmp_str(const char* a, int la, const char* b, int lb) {
    int m = la < lb ? la : lb;
    int c = memcmp(a, b, (size_t)m);
    if (c) return c;
    return la - lb;
}
static int allele_desc(const void* x, const void* y) {
    const allele_t* a = (const allele_t*)x; const allele_t* b = (const allele_t*)y;
    if (a->count != b->count) return a->count > b->count ? -1 : 1;
    if (a->freq != b->freq) return a->freq > b->freq ? -1 : 1;
    int c = cmp_str(a->s, a->len, b->s, b->len);
    return c > 0 ? -1 : (c < 0 ? 1 : 0);
}
static const char* FIXED_SYMS = "ACGTN-";

int64_t oracle_call(int L, const int64_t* counts, int64_t K, const int32_t* ins_pos, const int64_t* ins_count,
                    const int64_t* ins_off, const char* ins_chars, const char* ref_seq,
                    int run_consensus, int64_t min_depth_consensus, double min_freq_consensus,
                    int run_variants, int64_t min_depth_variants, double min_freq_variants,
                    int64_t* depth, int64_t* al_off, int64_t* al_count, double* al_freq, int32_t* al_sym, uint8_t* al_is_alt,
                    int32_t* cons_sym, uint8_t* var_emit, uint8_t* var_gt_ref, int64_t* var_ref_count, double* var_ref_freq) {
    int64_t k = 0, na_total = 0;
    allele_t* al = NULL; int64_t cap = 0;
    for (int p = 0; p < L; ++p) {
        int64_t k0 = k; while (k < K && ins_pos[k] == p) k++;
        int64_t n_here = 6 + (k - k0);
        if (n_here > cap) { cap = 2 * n_here; al = (allele_t*)realloc(al, sizeof(allele_t) * cap); }
        int64_t total = 0; int na = 0;                                          /* 767 */
        for (int ch = 0; ch < 6; ++ch) total += counts[ch * (int64_t)L + p];
        for (int64_t j = k0; j < k; ++j) total += ins_count[j];
        depth[p] = total; al_off[p] = na_total; cons_sym[p] = -1;
        var_emit[p] = 0; var_gt_ref[p] = 0; var_ref_count[p] = 0; var_ref_freq[p] = 0.0;
        if (total == 0) continue;                                               /* 768-769 */
        for (int ch = 0; ch < 6; ++ch) { int64_t cnt = counts[ch * (int64_t)L + p]; if (cnt != 0) {
            al[na].count = cnt; al[na].freq = (double)cnt / (double)total; al[na].s = FIXED_SYMS + ch; al[na].len = 1; al[na].id = ch; na++; } }
        for (int64_t j = k0; j < k; ++j) if (ins_count[j] != 0) {
            al[na].count = ins_count[j]; al[na].freq = (double)ins_count[j] / (double)total;
            al[na].s = ins_chars + ins_off[j]; al[na].len = (int)(ins_off[j + 1] - ins_off[j]); al[na].id = 6 + (int)j; na++; }
        qsort(al, (size_t)na, sizeof(allele_t), allele_desc);                   /* 771 */
        if (run_consensus && na != 0 && al[0].count >= min_depth_consensus && al[0].freq >= min_freq_consensus)
            cons_sym[p] = al[0].id;                                             /* 928-929 */
        int64_t tot_count = 0, ref_count = 0; double ref_freq = 0.0; int n_alt = 0;
        for (int a = 0; a < na; ++a) {                                          /* 934-939 */
            al_count[na_total + a] = al[a].count; al_freq[na_total + a] = al[a].freq; al_sym[na_total + a] = al[a].id; al_is_alt[na_total + a] = 0;
            tot_count += al[a].count;
            if (al[a].len == 1 && al[a].s[0] == ref_seq[p]) { ref_count = al[a].count; ref_freq = al[a].freq; }
            else if (al[a].freq >= min_freq_variants) { al_is_alt[na_total + a] = 1; n_alt++; }
        }
        if (run_variants && tot_count >= min_depth_variants && n_alt != 0) {    /* 940 */
            var_emit[p] = 1; var_ref_count[p] = ref_count; var_ref_freq[p] = ref_freq;
            var_gt_ref[p] = (ref_count >= min_depth_variants && ref_freq >= min_freq_variants) ? 1 : 0;   /* 948 */
        }
        na_total += na;
    }
    al_off[L] = na_total;
    free(al);
    return na_total;
}
