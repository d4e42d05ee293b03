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
} Job;

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
static void run_jobs(const Handle *h, void *(*fn)(void *), Job *jobs, int T) {
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
    for (Py_ssize_t k = 0; k < n; ++k) {
        PyObject *s = PyList_GET_ITEM(arg, k);
        if (!PyUnicode_Check(s) || !PyUnicode_IS_COMPACT_ASCII(s)) {       /* not ASCII: general path */
            free(h->ptr); free(h->off); free(h);
            Py_RETURN_NONE;
        }
        h->ptr[k] = (const unsigned char *)PyUnicode_1BYTE_DATA(s);
        h->off[k + 1] = h->off[k] + (int64_t)PyUnicode_GET_LENGTH(s);
    }
    Py_INCREF(arg);
    h->list = arg;
    Job jobs[16];
    const int T = n_threads(n);
    run_jobs(h, scan_job, jobs, T);
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
    run_jobs(h, translate_job, jobs, T);
    PyBuffer_Release(&table);
    return out;
}

static PyMethodDef methods[] = {
    {"scan", py_scan, METH_O, "scan(list[str]) -> (handle, total, lengths int64 bytes, seen[256]) | None"},
    {"translate", py_translate, METH_VARARGS, "translate(handle, table256) -> bytes"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_strings", "string-list lowering for globalign_b200", -1,
                                    methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__strings(void) { return PyModule_Create(&moddef); }
