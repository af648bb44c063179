/* _hostfast -- CPython helpers for the host-side ingestion of one tick.
 *
 * The reference builds one KubePod per kube-API dict in Python (kube.py:23-49) and then walks the pod
 * list several times.  At 10^5 pods the interpreter overhead of that construction is the largest host
 * cost of a tick, so the two hot loops have a C twin here:
 *
 *   make_pods(cls, raw_pods, time_memo, remember_time, res_memo, pod_resources) -> list
 *       KubePod.__init__ (kube.py of this package) for every element of raw_pods, attribute for
 *       attribute and in the same order.  ONLY the plain case is handled in C (exact dicts, exact str
 *       timestamps, a one-container spec whose requests hit the memo); anything else -- a missing key,
 *       an unexpected type, a memo miss -- is delegated to the Python code (cls(pod), remember_time(text),
 *       pod_resources(containers)), so errors and odd inputs behave exactly as in the Python class.
 *
 *   group_ids(seq, out) -> list
 *       groups the objects of seq by identity in first-occurrence order: out[i] (int64 buffer, len(seq))
 *       is the group of seq[i], the returned list holds one representative per group (snapshot.Dims).
 *
 * This is host logic only: no resource arithmetic happens here (that is the CUDA library's job).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static PyObject *k_obj, *k_metadata, *k_spec, *k_status, *k_name, *k_namespace, *k_nodeName, *k_phase, *k_uid,
    *k_nodeSelector, *k_labels, *k_annotations, *k_owner, *k_creationTimestamp, *k_startTime, *k_containers,
    *k_resources, *k_requests;
static PyObject *a_original, *a_name, *a_namespace, *a_node_name, *a_status, *a_uid, *a_selectors, *a_labels,
    *a_annotations, *a_owner, *a_creation_time, *a_start_time, *a_resources;
static PyObject *empty_tuple;

/* dict[key] for an exact dict; NULL without an exception when the key is missing */
static inline PyObject *get(PyObject *d, PyObject *key) { return PyDict_GetItemWithError(d, key); }

/* timestamp text -> parsed value: the memo when the text is an exact str and cached, else remember_time(text).
 * Returns a new reference, NULL with an exception set on failure. */
static PyObject *parse_time(PyObject *text, PyObject *time_memo, PyObject *remember_time)
{
    if (PyUnicode_CheckExact(text)) {
        PyObject *hit = PyDict_GetItemWithError(time_memo, text);
        if (hit && hit != Py_None) {
            Py_INCREF(hit);
            return hit;
        }
        if (!hit && PyErr_Occurred()) return NULL;
    }
    return PyObject_CallOneArg(remember_time, text);
}

/* the memo key of kube._pod_resources for a one-container spec; returns a new reference, or NULL without an
 * exception when the shape is not the plain one (the caller then takes the Python route) */
static PyObject *single_container_key(PyObject *containers)
{
    if (!PyList_CheckExact(containers) || PyList_GET_SIZE(containers) != 1) return NULL;
    PyObject *c0 = PyList_GET_ITEM(containers, 0);
    if (!PyDict_CheckExact(c0)) return NULL;
    PyObject *r = get(c0, k_resources);
    if (!r) {
        if (PyErr_Occurred()) PyErr_Clear();
        Py_INCREF(empty_tuple);
        return empty_tuple;  /* no 'resources': r is None -> q is None -> key () */
    }
    if (r == Py_None) { Py_INCREF(empty_tuple); return empty_tuple; }
    if (!PyDict_CheckExact(r)) return NULL;
    if (PyDict_GET_SIZE(r) == 0) { Py_INCREF(empty_tuple); return empty_tuple; }  /* falsy: q = None */
    PyObject *q = get(r, k_requests);
    if (!q) {
        if (PyErr_Occurred()) PyErr_Clear();
        Py_INCREF(empty_tuple);
        return empty_tuple;
    }
    if (q == Py_None) { Py_INCREF(empty_tuple); return empty_tuple; }
    if (!PyDict_CheckExact(q)) return NULL;
    Py_ssize_t n = PyDict_GET_SIZE(q);
    if (n == 0) { Py_INCREF(empty_tuple); return empty_tuple; }
    PyObject *key = PyTuple_New(n);  /* tuple(q.items()) */
    if (!key) { PyErr_Clear(); return NULL; }
    Py_ssize_t pos = 0, i = 0;
    PyObject *k, *v;
    while (PyDict_Next(q, &pos, &k, &v)) {
        PyObject *item = PyTuple_Pack(2, k, v);
        if (!item) { PyErr_Clear(); Py_DECREF(key); return NULL; }
        PyTuple_SET_ITEM(key, i++, item);
    }
    return key;
}

static PyObject *make_pods(PyObject *self, PyObject *args)
{
    PyObject *cls, *raw, *time_memo, *remember_time, *res_memo, *pod_resources;
    if (!PyArg_ParseTuple(args, "OOO!OO!O", &cls, &raw, &PyDict_Type, &time_memo, &remember_time, &PyDict_Type,
                          &res_memo, &pod_resources))
        return NULL;
    if (!PyType_Check(cls)) {
        PyErr_SetString(PyExc_TypeError, "cls must be a class");
        return NULL;
    }
    PyObject *seq = PySequence_Fast(raw, "raw_pods must be iterable");
    if (!seq) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    PyObject *out = PyList_New(n);
    if (!out) { Py_DECREF(seq); return NULL; }
    PyTypeObject *tp = (PyTypeObject *)cls;

    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *pod = PySequence_Fast_GET_ITEM(seq, i);
        PyObject *inst = NULL, *obj = NULL, *selectors = NULL, *labels = NULL, *annotations = NULL, *ctime = NULL,
                 *stime = NULL, *res = NULL, *key = NULL;
        int plain = 0;
        obj = PyObject_GetAttr(pod, k_obj);
        if (!obj) { PyErr_Clear(); goto slow; }
        if (!PyDict_CheckExact(obj)) goto slow;
        {
            PyObject *meta = get(obj, k_metadata), *spec = get(obj, k_spec), *status = get(obj, k_status);
            if (!meta || !spec || !status || !PyDict_CheckExact(meta) || !PyDict_CheckExact(spec) ||
                !PyDict_CheckExact(status))
                goto slow;
            PyObject *name = get(meta, k_name), *ns = get(meta, k_namespace), *uid = get(meta, k_uid),
                     *phase = get(status, k_phase), *ctext = get(meta, k_creationTimestamp),
                     *containers = get(spec, k_containers);
            if (!name || !ns || !uid || !phase || !ctext || !containers) goto slow;
            PyObject *node_name = get(spec, k_nodeName);
            if (!node_name) { if (PyErr_Occurred()) goto slow; node_name = Py_None; }
            selectors = get(spec, k_nodeSelector);
            if (selectors) Py_INCREF(selectors); else { if (PyErr_Occurred()) goto slow; selectors = PyDict_New(); }
            labels = get(meta, k_labels);
            if (labels) Py_INCREF(labels); else { if (PyErr_Occurred()) goto slow; labels = PyDict_New(); }
            annotations = get(meta, k_annotations);
            if (annotations) Py_INCREF(annotations); else { if (PyErr_Occurred()) goto slow; annotations = PyDict_New(); }
            if (!selectors || !labels || !annotations) goto slow;
            if (!PyDict_CheckExact(labels)) goto slow;  /* labels.get('owner') on something else: Python decides */
            PyObject *owner = get(labels, k_owner);
            if (!owner) { if (PyErr_Occurred()) goto slow; owner = Py_None; }
            /* the resource memo first: a miss or an odd shape sends the whole pod down the Python route, so that a
             * failing constructor has had no side effect here */
            key = single_container_key(containers);
            if (!key) goto slow;
            res = PyDict_GetItemWithError(res_memo, key);
            if (!res) { if (PyErr_Occurred()) PyErr_Clear(); goto slow; }
            Py_INCREF(res);
            ctime = parse_time(ctext, time_memo, remember_time);
            if (!ctime) goto fail;  /* a malformed timestamp raises, as in the Python constructor */
            PyObject *stext = get(status, k_startTime);
            if (!stext) {
                if (PyErr_Occurred()) goto fail;
                stime = Py_None;
                Py_INCREF(stime);
            } else {
                stime = parse_time(stext, time_memo, remember_time);
                if (!stime) goto fail;
            }
            inst = tp->tp_new(tp, empty_tuple, NULL);
            if (!inst) goto fail;
            if (PyObject_SetAttr(inst, a_original, pod) < 0 || PyObject_SetAttr(inst, a_name, name) < 0 ||
                PyObject_SetAttr(inst, a_namespace, ns) < 0 || PyObject_SetAttr(inst, a_node_name, node_name) < 0 ||
                PyObject_SetAttr(inst, a_status, phase) < 0 || PyObject_SetAttr(inst, a_uid, uid) < 0 ||
                PyObject_SetAttr(inst, a_selectors, selectors) < 0 || PyObject_SetAttr(inst, a_labels, labels) < 0 ||
                PyObject_SetAttr(inst, a_annotations, annotations) < 0 || PyObject_SetAttr(inst, a_owner, owner) < 0 ||
                PyObject_SetAttr(inst, a_creation_time, ctime) < 0 || PyObject_SetAttr(inst, a_start_time, stime) < 0 ||
                PyObject_SetAttr(inst, a_resources, res) < 0)
                goto fail;
            plain = 1;
        }
    slow:
        if (!plain) {
            if (PyErr_Occurred()) PyErr_Clear();
            Py_XDECREF(inst);
            inst = PyObject_CallOneArg(cls, pod);  /* the Python constructor: same result, same exceptions */
            if (!inst) goto fail;
        }
        Py_XDECREF(obj); Py_XDECREF(selectors); Py_XDECREF(labels); Py_XDECREF(annotations);
        Py_XDECREF(ctime); Py_XDECREF(stime); Py_XDECREF(res); Py_XDECREF(key);
        PyList_SET_ITEM(out, i, inst);
        continue;
    fail:
        Py_XDECREF(inst); Py_XDECREF(obj); Py_XDECREF(selectors); Py_XDECREF(labels); Py_XDECREF(annotations);
        Py_XDECREF(ctime); Py_XDECREF(stime); Py_XDECREF(res); Py_XDECREF(key);
        Py_DECREF(out);
        Py_DECREF(seq);
        return NULL;
    }
    Py_DECREF(seq);
    return out;
}

/* identity grouping: open addressing on the object address */
static PyObject *group_ids(PyObject *self, PyObject *args)
{
    PyObject *raw;
    Py_buffer buf;
    if (!PyArg_ParseTuple(args, "Ow*", &raw, &buf)) return NULL;
    PyObject *seq = PySequence_Fast(raw, "seq must be iterable");
    if (!seq) { PyBuffer_Release(&buf); return NULL; }
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    if (buf.len < (Py_ssize_t)(n * sizeof(int64_t))) {
        PyBuffer_Release(&buf);
        Py_DECREF(seq);
        PyErr_SetString(PyExc_ValueError, "out must hold len(seq) int64 values");
        return NULL;
    }
    int64_t *out = (int64_t *)buf.buf;
    size_t cap = 1024;
    PyObject **keys = (PyObject **)calloc(cap, sizeof(PyObject *));
    int64_t *vals = (int64_t *)malloc(cap * sizeof(int64_t));
    PyObject *uniq = PyList_New(0);
    if (!keys || !vals || !uniq) goto nomem;
    size_t used = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *o = PySequence_Fast_GET_ITEM(seq, i);
        size_t h = ((uintptr_t)o >> 4) * 0x9E3779B97F4A7C15ull;
        size_t j = (h >> 20) & (cap - 1);
        while (keys[j] && keys[j] != o) j = (j + 1) & (cap - 1);
        if (!keys[j]) {
            keys[j] = o;
            vals[j] = (int64_t)used++;
            if (PyList_Append(uniq, o) < 0) goto nomem;
            out[i] = vals[j];
            if (used * 2 > cap) {  /* grow and rehash */
                size_t ncap = cap * 4;
                PyObject **nk = (PyObject **)calloc(ncap, sizeof(PyObject *));
                int64_t *nv = (int64_t *)malloc(ncap * sizeof(int64_t));
                if (!nk || !nv) { free(nk); free(nv); goto nomem; }
                for (size_t t = 0; t < cap; ++t)
                    if (keys[t]) {
                        size_t hh = ((uintptr_t)keys[t] >> 4) * 0x9E3779B97F4A7C15ull;
                        size_t jj = (hh >> 20) & (ncap - 1);
                        while (nk[jj]) jj = (jj + 1) & (ncap - 1);
                        nk[jj] = keys[t];
                        nv[jj] = vals[t];
                    }
                free(keys); free(vals);
                keys = nk; vals = nv; cap = ncap;
            }
        } else {
            out[i] = vals[j];
        }
    }
    free(keys); free(vals);
    PyBuffer_Release(&buf);
    Py_DECREF(seq);
    return uniq;
nomem:
    free(keys); free(vals);
    Py_XDECREF(uniq);
    PyBuffer_Release(&buf);
    Py_DECREF(seq);
    if (!PyErr_Occurred()) PyErr_NoMemory();
    return NULL;
}

static PyMethodDef methods[] = {
    {"make_pods", make_pods, METH_VARARGS, "KubePod construction for a list of kube-API pod objects"},
    {"group_ids", group_ids, METH_VARARGS, "group objects by identity in first-occurrence order"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_hostfast", "host-side ingestion helpers", -1, methods};

#define INTERN(var, text) do { var = PyUnicode_InternFromString(text); if (!var) return NULL; } while (0)

PyMODINIT_FUNC PyInit__hostfast(void)
{
    INTERN(k_obj, "obj"); INTERN(k_metadata, "metadata"); INTERN(k_spec, "spec"); INTERN(k_status, "status");
    INTERN(k_name, "name"); INTERN(k_namespace, "namespace"); INTERN(k_nodeName, "nodeName"); INTERN(k_phase, "phase");
    INTERN(k_uid, "uid"); INTERN(k_nodeSelector, "nodeSelector"); INTERN(k_labels, "labels");
    INTERN(k_annotations, "annotations"); INTERN(k_owner, "owner"); INTERN(k_creationTimestamp, "creationTimestamp");
    INTERN(k_startTime, "startTime"); INTERN(k_containers, "containers"); INTERN(k_resources, "resources");
    INTERN(k_requests, "requests");
    INTERN(a_original, "original"); INTERN(a_name, "name"); INTERN(a_namespace, "namespace");
    INTERN(a_node_name, "node_name"); INTERN(a_status, "status"); INTERN(a_uid, "uid"); INTERN(a_selectors, "selectors");
    INTERN(a_labels, "labels"); INTERN(a_annotations, "annotations"); INTERN(a_owner, "owner");
    INTERN(a_creation_time, "creation_time"); INTERN(a_start_time, "start_time"); INTERN(a_resources, "resources");
    empty_tuple = PyTuple_New(0);
    if (!empty_tuple) return NULL;
    return PyModule_Create(&moduledef);
}
