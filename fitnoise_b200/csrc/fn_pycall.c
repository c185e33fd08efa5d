/* Thin CPython entry for the one call that sits inside the optimiser loop.
 *
 * fn_nll_grad is called 20-60 times per fit and an evaluation of a small or strongly sharded matrix
 * takes 10-25 us on the GPU; the generic ctypes marshalling of its seven arguments (plus the numpy
 * copies around it) costs 2-3 us of host time per call.  This module calls the SAME C-ABI function
 * through its address (no link dependency: libfitnoise_b200.so stays a plain C library with no
 * Python symbols) with the arguments taken straight from the buffer protocol: ~0.3 us.
 * It computes nothing; without it (no Python.h / gcc where the package is built) _native.py uses ctypes.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

typedef int (*fn_nll_grad_t)(void*, const double*, int32_t, double*, double*, int64_t*, int64_t*);

/* nll_grad(function address, handle address, theta (float64 buffer), grad_out (writable float64 buffer))
 *   -> (rc, value, n_used, n_bad) */
static PyObject* nll_grad(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
    (void)self;
    if (nargs != 4) { PyErr_SetString(PyExc_TypeError, "nll_grad(fn, handle, theta, grad_out)"); return NULL; }
    fn_nll_grad_t fn = (fn_nll_grad_t)PyLong_AsVoidPtr(args[0]);
    void* handle = PyLong_AsVoidPtr(args[1]);
    if (PyErr_Occurred()) return NULL;
    if (!fn || !handle) { PyErr_SetString(PyExc_ValueError, "nll_grad: null function or handle address"); return NULL; }
    Py_buffer theta, grad;
    if (PyObject_GetBuffer(args[2], &theta, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) return NULL;
    if (PyObject_GetBuffer(args[3], &grad, PyBUF_C_CONTIGUOUS | PyBUF_WRITABLE | PyBUF_FORMAT) != 0) {
        PyBuffer_Release(&theta);
        return NULL;
    }
    PyObject* result = NULL;
    const int ok = theta.itemsize == 8 && grad.itemsize == 8 && theta.format && grad.format &&
                   theta.format[0] == 'd' && grad.format[0] == 'd' && grad.len >= theta.len;
    if (!ok) {
        PyErr_SetString(PyExc_TypeError, "theta and grad_out must be contiguous float64 buffers, grad_out at least as long");
    } else {
        double value = 0.0;
        int64_t used = 0, bad = 0;
        int rc;
        const int32_t P = (int32_t)(theta.len / 8);
        Py_BEGIN_ALLOW_THREADS
        rc = fn(handle, (const double*)theta.buf, P, &value, (double*)grad.buf, &used, &bad);
        Py_END_ALLOW_THREADS
        result = Py_BuildValue("(idLL)", rc, value, (long long)used, (long long)bad);
    }
    PyBuffer_Release(&theta);
    PyBuffer_Release(&grad);
    return result;
}

static PyMethodDef methods[] = {
    {"nll_grad", (PyCFunction)(void (*)(void))nll_grad, METH_FASTCALL, "fn_nll_grad through its address"},
    {NULL, NULL, 0, NULL},
};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_fn_pycall", "fast path into fn_nll_grad", -1, methods,
                                    NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__fn_pycall(void) { return PyModule_Create(&module); }
