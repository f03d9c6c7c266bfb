/* CPython helper of yastn_b200.plans: flatten YASTN's nested meta tuples (tuples / lists of Python ints) into one
 * little int64 buffer in depth-first order.  The reference hands the backend its block metadata as nested tuples
 * (yastn/tensor/_merging.py:84-89,137-187,528-549; _contractions.py:281-346); turning thousands of records into the
 * int64 tables of the C ABI with Python-level loops costs more than the kernel launch they describe.
 *
 *     flatten(obj) -> bytes        (np.frombuffer(..., dtype=np.int64) on the Python side)
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
    int64_t* data;
    Py_ssize_t n, cap;
} Buf;

static int push(Buf* b, int64_t v) {
    if (b->n == b->cap) {
        Py_ssize_t cap = b->cap ? b->cap * 2 : 1024;
        int64_t* p = (int64_t*)realloc(b->data, (size_t)cap * sizeof(int64_t));
        if (!p) {
            PyErr_NoMemory();
            return -1;
        }
        b->data = p;
        b->cap = cap;
    }
    b->data[b->n++] = v;
    return 0;
}

static int walk(PyObject* o, Buf* b, int depth) {
    if (PyLong_Check(o)) {
        int overflow = 0;
        long long v = PyLong_AsLongLongAndOverflow(o, &overflow);
        if (overflow || (v == -1 && PyErr_Occurred())) {
            if (!PyErr_Occurred()) PyErr_SetString(PyExc_OverflowError, "flatten: integer does not fit in int64");
            return -1;
        }
        return push(b, (int64_t)v);
    }
    if (PyFloat_Check(o)) {   /* the reference's metas carry np.prod(()) == 1.0 for empty leg groups (outer products) */
        double d = PyFloat_AS_DOUBLE(o);
        if (d != (double)(int64_t)d) {
            PyErr_SetString(PyExc_ValueError, "flatten: non-integral float in meta");
            return -1;
        }
        return push(b, (int64_t)d);
    }
    if (depth > 64) {
        PyErr_SetString(PyExc_ValueError, "flatten: nesting too deep");
        return -1;
    }
    if (PyTuple_Check(o)) {
        Py_ssize_t n = PyTuple_GET_SIZE(o);
        for (Py_ssize_t i = 0; i < n; ++i)
            if (walk(PyTuple_GET_ITEM(o, i), b, depth + 1)) return -1;
        return 0;
    }
    if (PyList_Check(o)) {
        Py_ssize_t n = PyList_GET_SIZE(o);
        for (Py_ssize_t i = 0; i < n; ++i)
            if (walk(PyList_GET_ITEM(o, i), b, depth + 1)) return -1;
        return 0;
    }
    {   /* numpy integers and other index-like scalars */
        PyObject* idx = PyNumber_Index(o);
        if (!idx) {
            PyErr_Format(PyExc_TypeError, "flatten: unsupported element of type %s", Py_TYPE(o)->tp_name);
            return -1;
        }
        int rc = walk(idx, b, depth + 1);
        Py_DECREF(idx);
        return rc;
    }
}

static PyObject* flatten(PyObject* self, PyObject* arg) {
    Buf b = {NULL, 0, 0};
    if (walk(arg, &b, 0)) {
        free(b.data);
        return NULL;
    }
    PyObject* out = PyBytes_FromStringAndSize((const char*)b.data, b.n * (Py_ssize_t)sizeof(int64_t));
    free(b.data);
    return out;
}

static PyMethodDef methods[] = {{"flatten", flatten, METH_O, "Flatten nested tuples/lists of ints into int64 bytes (depth-first)."},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_flatten", "meta tuple flattener", -1, methods};
PyMODINIT_FUNC PyInit__flatten(void) { return PyModule_Create(&moddef); }
