/* _strings: lowering of Python string lists for the batched entry point (CPython extension).
 *
 * The reference takes two str per call and upper-cases them (start.py:219-220); its dict-of-dict
 * costing matrix is indexed by the characters.  find_global_alignments takes lists of a million
 * strings; turning them into the letter indices the C-ABI consumes with pure Python / numpy
 * (join, encode, translate) costs ~0.9 s per million 150-letter pairs -- 90x the GPU time.  This
 * module does the same work in two parallel passes over the strings' own buffers:
 *
 *   scan(seqs)            -> handle, total letters, int64 lengths, 256 "byte value occurs" flags,
 *                            or None when a sequence is not a compact ASCII str (the caller then
 *                            takes the general Python path)
 *   translate(handle, t)  -> bytes: every letter of every sequence mapped through the 256-entry
 *                            table t (case folding and letter -> index in one pass), concatenated
 *   translate_into(handle, t, buffer, bits)
 *                         -> None: the same into a writable buffer the caller owns (pinned host
 *                            memory: no page faults, and the upload is a plain DMA), one byte per
 *                            letter (bits = 8) or as the C-ABI's 2- / 4- / 5-bit little-endian
 *                            letter streams (GA_F_SEQ_2BIT / _4BIT / _5BIT, include/globalign_b200.h)
 *
 * The strings of a million read pairs are two million separate heap objects: even reading their
 * headers (type, kind, length, data pointer) is a cache miss each, so that pass is threaded too.
 *
 * The GIL is held throughout (the worker threads touch no Python object, only the character
 * buffers the handle points into; the handle keeps the list alive).  No alignment work happens
 * here; there is no CPU fallback of the DP in this package.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    PyObject *list;          /* borrowed buffers stay valid while this reference lives */
    Py_ssize_t n;
    const unsigned char **ptr;
    int64_t *off;            /* n + 1 */
} Handle;

static void handle_free(PyObject *cap) {
    Handle *h = (Handle *)PyCapsule_GetPointer(cap, "ga_strings.handle");
    if (!h) return;
    Py_XDECREF(h->list);
    free(h->ptr);
    free(h->off);
    free(h);
}

typedef struct {
    const Handle *h;
    Py_ssize_t lo, hi;
    unsigned char seen[256];
    const unsigned char *table;
    unsigned char *out;
    PyObject **items;        /* header pass: the list's items */
    int64_t *lens;
    int bad;                 /* header pass: an item is not a compact ASCII str */
    int bits;                /* translate_into: 8, 5, 4 or 2 */
} Job;

/* type, kind, length and data pointer of every string of [lo, hi): reads object headers only (no
 * reference counts, no allocation); the calling thread holds the GIL throughout, so nothing can
 * change them */
static void *header_job(void *arg) {
    Job *j = (Job *)arg;
    Handle *h = (Handle *)j->h;
    j->bad = 0;
    for (Py_ssize_t k = j->lo; k < j->hi; ++k) {
        PyObject *s = j->items[k];
        if (!PyUnicode_Check(s) || !PyUnicode_IS_COMPACT_ASCII(s)) { j->bad = 1; return NULL; }
        h->ptr[k] = (const unsigned char *)PyUnicode_1BYTE_DATA(s);
        j->lens[k] = (int64_t)PyUnicode_GET_LENGTH(s);
    }
    return NULL;
}

/* my sequences as a little-endian stream of `bits`-wide fields (letter k of the concatenation in
 * bits [k*bits, (k+1)*bits) of the output).  A byte that two sequences share is written first by
 * the tail of the earlier one (assignment, upper fields zero) and then OR-ed into by the head of
 * the later one; jobs start on byte boundaries (run_jobs), so both happen in one thread. */
static void *pack_job(void *arg) {
    Job *j = (Job *)arg;
    const int bits = j->bits, per = 8 / bits;
    const unsigned char *t = j->table;
    for (Py_ssize_t k = j->lo; k < j->hi; ++k) {
        const unsigned char *p = j->h->ptr[k];
        int64_t at = j->h->off[k];
        const int64_t end = j->h->off[k + 1];
        while (at < end && at % per) {                          /* head */
            j->out[at / per] |= (unsigned char)(t[*p++] << ((at % per) * bits));
            ++at;
        }
        unsigned char *o = j->out + at / per;
        if (bits == 2) {
            for (; at + 4 <= end; at += 4, p += 4)
                *o++ = (unsigned char)(t[p[0]] | (t[p[1]] << 2) | (t[p[2]] << 4) | (t[p[3]] << 6));
        } else {
            for (; at + 2 <= end; at += 2, p += 2) *o++ = (unsigned char)(t[p[0]] | (t[p[1]] << 4));
        }
        if (at < end) {                                         /* tail */
            unsigned v = 0;
            for (int f = 0; at < end; ++at, ++f) v |= (unsigned)t[*p++] << (f * bits);
            *o = (unsigned char)v;
        }
    }
    return NULL;
}

/* 5-bit fields (amino acids): letter k in bits [5k, 5k + 5), so a letter may straddle two bytes and
 * eight letters fill five.  Jobs start on multiples of eight letters (run_jobs), i.e. on a byte
 * boundary; each job zeroes its own bytes first and ORs the letters in. */
static void *pack5_job(void *arg) {
    Job *j = (Job *)arg;
    const unsigned char *t = j->table;
    if (j->lo >= j->hi) return NULL;
    {
        const int64_t b0 = j->h->off[j->lo] * 5 / 8, b1 = (j->h->off[j->hi] * 5 + 7) / 8;
        memset(j->out + b0, 0, (size_t)(b1 - b0));
    }
    for (Py_ssize_t k = j->lo; k < j->hi; ++k) {
        const unsigned char *p = j->h->ptr[k];
        int64_t at = j->h->off[k];
        const int64_t end = j->h->off[k + 1];
        for (; at < end && (at & 7); ++at) {                    /* head: up to the next 5-byte group */
            const int64_t bit = at * 5;
            const unsigned v = (unsigned)t[*p++] << (bit & 7);
            j->out[bit >> 3] |= (unsigned char)v;
            if (v >> 8) j->out[(bit >> 3) + 1] |= (unsigned char)(v >> 8);
        }
        for (; at + 8 <= end; at += 8, p += 8) {                /* whole groups: 8 letters -> 5 bytes */
            uint64_t w = 0;
            for (int f = 0; f < 8; ++f) w |= (uint64_t)t[p[f]] << (5 * f);
            unsigned char *o = j->out + at * 5 / 8;
            o[0] = (unsigned char)w; o[1] = (unsigned char)(w >> 8); o[2] = (unsigned char)(w >> 16);
            o[3] = (unsigned char)(w >> 24); o[4] = (unsigned char)(w >> 32);
        }
        for (; at < end; ++at) {                                /* tail */
            const int64_t bit = at * 5;
            const unsigned v = (unsigned)t[*p++] << (bit & 7);
            j->out[bit >> 3] |= (unsigned char)v;
            if (v >> 8) j->out[(bit >> 3) + 1] |= (unsigned char)(v >> 8);
        }
    }
    return NULL;
}

static void *scan_job(void *arg) {
    Job *j = (Job *)arg;
    memset(j->seen, 0, sizeof j->seen);
    for (Py_ssize_t k = j->lo; k < j->hi; ++k) {
        const unsigned char *p = j->h->ptr[k];
        const int64_t len = j->h->off[k + 1] - j->h->off[k];
        for (int64_t i = 0; i < len; ++i) j->seen[p[i]] = 1;
    }
    return NULL;
}

static void *translate_job(void *arg) {
    Job *j = (Job *)arg;
    for (Py_ssize_t k = j->lo; k < j->hi; ++k) {
        const unsigned char *p = j->h->ptr[k];
        const int64_t len = j->h->off[k + 1] - j->h->off[k];
        unsigned char *o = j->out + j->h->off[k];
        for (int64_t i = 0; i < len; ++i) o[i] = j->table[p[i]];
    }
    return NULL;
}

static int n_threads(Py_ssize_t n) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    if (c < 1) c = 1;
    if (c > 16) c = 16;
    if (n < 4096) c = 1;
    return (int)c;
}

/* split [0, n) so that every thread gets about the same number of LETTERS */
static void run_jobs(const Handle *h, void *(*fn)(void *), Job *jobs, int T, int per) {
    pthread_t th[16];
    const int64_t total = h->off[h->n];
    Py_ssize_t at = 0;
    for (int t = 0; t < T; ++t) {
        jobs[t].h = h;
        jobs[t].lo = at;
        const int64_t want = total * (t + 1) / T;
        Py_ssize_t lo = at, hi = h->n;
        while (lo < hi) {                 /* first k with off[k] >= want */
            Py_ssize_t mid = (lo + hi) / 2;
            if (h->off[mid] >= want) hi = mid; else lo = mid + 1;
        }
        while (lo < h->n && h->off[lo] % per) ++lo;        /* packed output: jobs start on a byte boundary */
        at = t + 1 == T ? h->n : lo;
        jobs[t].hi = at;
    }
    for (int t = 1; t < T; ++t) pthread_create(&th[t], NULL, fn, &jobs[t]);
    fn(&jobs[0]);
    for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
}

static PyObject *py_scan(PyObject *self, PyObject *arg) {
    (void)self;
    if (!PyList_Check(arg)) { PyErr_SetString(PyExc_TypeError, "scan() takes a list of str"); return NULL; }
    const Py_ssize_t n = PyList_GET_SIZE(arg);
    Handle *h = (Handle *)calloc(1, sizeof(Handle));
    if (!h) return PyErr_NoMemory();
    h->ptr = (const unsigned char **)malloc((size_t)(n + 1) * sizeof(void *));
    h->off = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
    if (!h->ptr || !h->off) { free(h->ptr); free(h->off); free(h); return PyErr_NoMemory(); }
    h->n = n;
    h->off[0] = 0;
    Job jobs[16];
    const int T = n_threads(n);
    {   /* headers, threaded over equal slices of the list; lengths land in off[1..] and are summed after */
        pthread_t th[16];
        PyObject **items = ((PyListObject *)arg)->ob_item;
        for (int t = 0; t < T; ++t) {
            jobs[t].h = h;
            jobs[t].items = items;
            jobs[t].lens = h->off + 1;
            jobs[t].lo = n * t / T;
            jobs[t].hi = n * (t + 1) / T;
        }
        for (int t = 1; t < T; ++t) pthread_create(&th[t], NULL, header_job, &jobs[t]);
        header_job(&jobs[0]);
        for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
        int bad = 0;
        for (int t = 0; t < T; ++t) bad |= jobs[t].bad;
        if (bad) {                                                          /* not ASCII: general path */
            free(h->ptr); free(h->off); free(h);
            Py_RETURN_NONE;
        }
        for (Py_ssize_t k = 0; k < n; ++k) h->off[k + 1] += h->off[k];
    }
    Py_INCREF(arg);
    h->list = arg;
    run_jobs(h, scan_job, jobs, T, 1);
    unsigned char seen[256];
    memset(seen, 0, sizeof seen);
    for (int t = 0; t < T; ++t)
        for (int c = 0; c < 256; ++c) seen[c] |= jobs[t].seen[c];
    PyObject *lens = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)n * 8);
    if (!lens) { free(h->ptr); free(h->off); Py_DECREF(arg); free(h); return NULL; }
    int64_t *lp = (int64_t *)PyBytes_AS_STRING(lens);
    for (Py_ssize_t k = 0; k < n; ++k) lp[k] = h->off[k + 1] - h->off[k];
    PyObject *cap = PyCapsule_New(h, "ga_strings.handle", handle_free);
    if (!cap) { Py_DECREF(lens); free(h->ptr); free(h->off); Py_DECREF(arg); free(h); return NULL; }
    return Py_BuildValue("(NLNy#)", cap, (long long)h->off[n], lens, (const char *)seen, (Py_ssize_t)256);
}

static PyObject *py_translate(PyObject *self, PyObject *args) {
    (void)self;
    PyObject *cap;
    Py_buffer table;
    if (!PyArg_ParseTuple(args, "Oy*", &cap, &table)) return NULL;
    Handle *h = (Handle *)PyCapsule_GetPointer(cap, "ga_strings.handle");
    if (!h || table.len != 256) {
        PyBuffer_Release(&table);
        if (h) PyErr_SetString(PyExc_ValueError, "translate() needs a 256-byte table");
        return NULL;
    }
    PyObject *out = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)h->off[h->n]);
    if (!out) { PyBuffer_Release(&table); return NULL; }
    Job jobs[16];
    const int T = n_threads(h->n);
    for (int t = 0; t < T; ++t) {
        jobs[t].table = (const unsigned char *)table.buf;
        jobs[t].out = (unsigned char *)PyBytes_AS_STRING(out);
    }
    run_jobs(h, translate_job, jobs, T, 1);
    PyBuffer_Release(&table);
    return out;
}

static PyObject *py_translate_into(PyObject *self, PyObject *args) {
    (void)self;
    PyObject *cap;
    Py_buffer table, out;
    int bits = 8;
    if (!PyArg_ParseTuple(args, "Oy*w*|i", &cap, &table, &out, &bits)) return NULL;
    Handle *h = (Handle *)PyCapsule_GetPointer(cap, "ga_strings.handle");
    const char *err = NULL;
    if (!h) { PyBuffer_Release(&table); PyBuffer_Release(&out); return NULL; }
    if (table.len != 256) err = "translate_into() needs a 256-byte table";
    else if (bits != 8 && bits != 5 && bits != 4 && bits != 2) err = "bits must be 8, 5, 4 or 2";
    else if ((int64_t)out.len < (h->off[h->n] * bits + 7) / 8) err = "output buffer too small";
    if (err) {
        PyBuffer_Release(&table); PyBuffer_Release(&out);
        PyErr_SetString(PyExc_ValueError, err);
        return NULL;
    }
    Job jobs[16];
    const int T = n_threads(h->n);
    for (int t = 0; t < T; ++t) {
        jobs[t].table = (const unsigned char *)table.buf;
        jobs[t].out = (unsigned char *)out.buf;
        jobs[t].bits = bits;
    }
    if (bits == 5) run_jobs(h, pack5_job, jobs, T, 8);
    else run_jobs(h, bits == 8 ? translate_job : pack_job, jobs, T, 8 / bits);
    PyBuffer_Release(&table);
    PyBuffer_Release(&out);
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"scan", py_scan, METH_O, "scan(list[str]) -> (handle, total, lengths int64 bytes, seen[256]) | None"},
    {"translate", py_translate, METH_VARARGS, "translate(handle, table256) -> bytes"},
    {"translate_into", py_translate_into, METH_VARARGS,
     "translate_into(handle, table256, writable buffer, bits=8): letters (bits = 8) or packed 4-/2-bit letter streams"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_strings", "string-list lowering for globalign_b200", -1,
                                    methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__strings(void) { return PyModule_Create(&moddef); }
