/*
 * Bulk construction of cv2.DMatch result objects (host side of the drop-in, not the data path).
 *
 * cv2's Python binding wraps cv::DMatch as { PyObject_HEAD; cv::DMatch v; } with
 * v = { int queryIdx; int trainIdx; int imgIdx; float distance; } (basicsize 32).  Calling the type
 * from Python costs 0.4-2.4 us per object (argument parsing), which dwarfs the GPU time of a
 * 2000 x 2000 match ("Python result materialisation").  Here objects are
 * allocated with the type's own tp_alloc and the four fields are written directly (~40 ns each).
 * The Python side verifies the layout once at import and falls back to the type call otherwise.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

typedef struct {
    PyObject_HEAD
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float distance;
} hm_dmatch_obj;

static PyObject* make_one(PyTypeObject* tp, int32_t q, int32_t t, int32_t img, float d)
{
    PyObject* o = tp->tp_alloc(tp, 0);
    if (!o) return NULL;
    hm_dmatch_obj* m = (hm_dmatch_obj*)o;
    m->queryIdx = q;
    m->trainIdx = t;
    m->imgIdx = img;
    m->distance = d;
    return o;
}

static int get_i32(PyObject* obj, Py_buffer* view, Py_ssize_t n, const char* name)
{
    if (PyObject_GetBuffer(obj, view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) return -1;
    if (view->itemsize != 4 || view->len != n * 4) {
        PyBuffer_Release(view);
        PyErr_Format(PyExc_ValueError, "%s must be a contiguous 4-byte array of %zd items", name, n);
        return -1;
    }
    return 0;
}

/* dmatch_list(type, q:int32[n], t:int32[n], d:float32[n], img:int32[n] or None, img_const:int, rows:int)
 * rows == 0 -> list of n DMatch; rows == k > 0 -> tuple of n/k tuples of k DMatch (knnMatch shape) */
static PyObject* dmatch_build(PyObject* self, PyObject* args)
{
    PyObject *tp_obj, *qo, *to, *dobj, *imgo;
    int img_const = 0, rows = 0;
    if (!PyArg_ParseTuple(args, "OOOOOii", &tp_obj, &qo, &to, &dobj, &imgo, &img_const, &rows)) return NULL;
    if (!PyType_Check(tp_obj) || ((PyTypeObject*)tp_obj)->tp_basicsize != (Py_ssize_t)sizeof(hm_dmatch_obj)) {
        PyErr_SetString(PyExc_TypeError, "not a DMatch-shaped type");
        return NULL;
    }
    PyTypeObject* tp = (PyTypeObject*)tp_obj;
    Py_buffer qv, tv, dv, iv;
    if (PyObject_GetBuffer(qo, &qv, PyBUF_C_CONTIGUOUS) < 0) return NULL;
    Py_ssize_t n = qv.len / 4;
    PyBuffer_Release(&qv);
    if (get_i32(qo, &qv, n, "q") < 0) return NULL;
    if (get_i32(to, &tv, n, "t") < 0) { PyBuffer_Release(&qv); return NULL; }
    if (get_i32(dobj, &dv, n, "d") < 0) { PyBuffer_Release(&qv); PyBuffer_Release(&tv); return NULL; }
    int have_img = imgo != Py_None;
    if (have_img && get_i32(imgo, &iv, n, "img") < 0) {
        PyBuffer_Release(&qv); PyBuffer_Release(&tv); PyBuffer_Release(&dv);
        return NULL;
    }
    const int32_t* q = (const int32_t*)qv.buf;
    const int32_t* t = (const int32_t*)tv.buf;
    const float* d = (const float*)dv.buf;
    const int32_t* im = have_img ? (const int32_t*)iv.buf : NULL;
    PyObject* out = NULL;
    if (rows <= 0) {
        out = PyList_New(n);
        for (Py_ssize_t i = 0; out && i < n; ++i) {
            PyObject* o = make_one(tp, q[i], t[i], im ? im[i] : img_const, d[i]);
            if (!o) { Py_CLEAR(out); break; }
            PyList_SET_ITEM(out, i, o);
        }
    } else if (n % rows == 0) {
        Py_ssize_t nr = n / rows;
        out = PyTuple_New(nr);
        for (Py_ssize_t r = 0; out && r < nr; ++r) {
            PyObject* row = PyTuple_New(rows);
            if (!row) { Py_CLEAR(out); break; }
            PyTuple_SET_ITEM(out, r, row);
            for (int j = 0; j < rows; ++j) {
                Py_ssize_t i = r * rows + j;
                PyObject* o = make_one(tp, q[i], t[i], im ? im[i] : img_const, d[i]);
                if (!o) { Py_CLEAR(out); break; }
                PyTuple_SET_ITEM(row, j, o);
            }
        }
    } else {
        PyErr_SetString(PyExc_ValueError, "n is not a multiple of rows");
    }
    PyBuffer_Release(&qv); PyBuffer_Release(&tv); PyBuffer_Release(&dv);
    if (have_img) PyBuffer_Release(&iv);
    return out;
}

/* Two-phase variant for calls whose result count is known before the GPU has finished (knnMatch: nq * k): the objects
 * and their containers are allocated while the kernels run, the fields are written after the device-to-host copy.
 * dmatch_alloc(type, n, rows) -> same containers as dmatch_build, every field zero */
static PyObject* dmatch_alloc(PyObject* self, PyObject* args)
{
    PyObject* tp_obj;
    Py_ssize_t n = 0;
    int rows = 0;
    if (!PyArg_ParseTuple(args, "Oni", &tp_obj, &n, &rows)) return NULL;
    if (!PyType_Check(tp_obj) || ((PyTypeObject*)tp_obj)->tp_basicsize != (Py_ssize_t)sizeof(hm_dmatch_obj) || n < 0 ||
        (rows > 0 && n % rows != 0)) {
        PyErr_SetString(PyExc_TypeError, "dmatch_alloc: not a DMatch-shaped type, or n is not a multiple of rows");
        return NULL;
    }
    PyTypeObject* tp = (PyTypeObject*)tp_obj;
    PyObject* out = NULL;
    if (rows <= 0) {
        out = PyList_New(n);
        for (Py_ssize_t i = 0; out && i < n; ++i) {
            PyObject* o = make_one(tp, 0, 0, 0, 0.0f);
            if (!o) { Py_CLEAR(out); break; }
            PyList_SET_ITEM(out, i, o);
        }
    } else {
        Py_ssize_t nr = n / rows;
        out = PyTuple_New(nr);
        for (Py_ssize_t r = 0; out && r < nr; ++r) {
            PyObject* row = PyTuple_New(rows);
            if (!row) { Py_CLEAR(out); break; }
            PyTuple_SET_ITEM(out, r, row);
            for (int j = 0; j < rows; ++j) {
                PyObject* o = make_one(tp, 0, 0, 0, 0.0f);
                if (!o) { Py_CLEAR(out); break; }
                PyTuple_SET_ITEM(row, j, o);
            }
        }
    }
    return out;
}

/* dmatch_fill(container, q, t, d, img or None, img_const, rows): writes the fields of the objects dmatch_alloc made */
static PyObject* dmatch_fill(PyObject* self, PyObject* args)
{
    PyObject *cont, *qo, *to, *dobj, *imgo;
    int img_const = 0, rows = 0;
    if (!PyArg_ParseTuple(args, "OOOOOii", &cont, &qo, &to, &dobj, &imgo, &img_const, &rows)) return NULL;
    Py_buffer qv, tv, dv, iv;
    if (PyObject_GetBuffer(qo, &qv, PyBUF_C_CONTIGUOUS) < 0) return NULL;
    Py_ssize_t n = qv.len / 4;
    PyBuffer_Release(&qv);
    if (get_i32(qo, &qv, n, "q") < 0) return NULL;
    if (get_i32(to, &tv, n, "t") < 0) { PyBuffer_Release(&qv); return NULL; }
    if (get_i32(dobj, &dv, n, "d") < 0) { PyBuffer_Release(&qv); PyBuffer_Release(&tv); return NULL; }
    int have_img = imgo != Py_None;
    if (have_img && get_i32(imgo, &iv, n, "img") < 0) {
        PyBuffer_Release(&qv); PyBuffer_Release(&tv); PyBuffer_Release(&dv);
        return NULL;
    }
    const int32_t* q = (const int32_t*)qv.buf;
    const int32_t* t = (const int32_t*)tv.buf;
    const float* d = (const float*)dv.buf;
    const int32_t* im = have_img ? (const int32_t*)iv.buf : NULL;
    int ok = 1;
    if (rows <= 0) {
        ok = PyList_Check(cont) && PyList_GET_SIZE(cont) == n;
        for (Py_ssize_t i = 0; ok && i < n; ++i) {
            hm_dmatch_obj* m = (hm_dmatch_obj*)PyList_GET_ITEM(cont, i);
            m->queryIdx = q[i]; m->trainIdx = t[i]; m->imgIdx = im ? im[i] : img_const; m->distance = d[i];
        }
    } else {
        ok = PyTuple_Check(cont) && n % rows == 0 && PyTuple_GET_SIZE(cont) == n / rows;
        for (Py_ssize_t r = 0; ok && r < n / rows; ++r) {
            PyObject* row = PyTuple_GET_ITEM(cont, r);
            if (!PyTuple_Check(row) || PyTuple_GET_SIZE(row) != rows) { ok = 0; break; }
            for (int j = 0; j < rows; ++j) {
                Py_ssize_t i = r * rows + j;
                hm_dmatch_obj* m = (hm_dmatch_obj*)PyTuple_GET_ITEM(row, j);
                m->queryIdx = q[i]; m->trainIdx = t[i]; m->imgIdx = im ? im[i] : img_const; m->distance = d[i];
            }
        }
    }
    PyBuffer_Release(&qv); PyBuffer_Release(&tv); PyBuffer_Release(&dv);
    if (have_img) PyBuffer_Release(&iv);
    if (!ok) {
        PyErr_SetString(PyExc_ValueError, "dmatch_fill: container does not match the arrays");
        return NULL;
    }
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"dmatch_alloc", dmatch_alloc, METH_VARARGS, "allocate DMatch containers (fields zero) ahead of the results"},
    {"dmatch_fill", dmatch_fill, METH_VARARGS, "write the fields of pre-allocated DMatch objects"},
    {"dmatch_build", dmatch_build, METH_VARARGS, "bulk-build DMatch objects from int32/float32 arrays"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_hmfast", NULL, -1, methods};

PyMODINIT_FUNC PyInit__hmfast(void) { return PyModule_Create(&moddef); }
