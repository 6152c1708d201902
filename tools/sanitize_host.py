"""Host code of libcldrd (run-file writer / reader / score formatter, index-file writer) under AddressSanitizer and
UndefinedBehaviorSanitizer.  CPU only:

    g++ -O1 -g -fsanitize=address,undefined -std=c++17 -Iinclude -Icl-drd_b200/csrc -shared -fPIC \
        cl-drd_b200/csrc/runfile.cpp cl-drd_b200/csrc/index_io.cpp -o /tmp/libhost_asan.so
    LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 \
        python tools/sanitize_host.py
"""
import ctypes as C, numpy as np, os, tempfile
L = C.CDLL("/tmp/libhost_asan.so")
L.cldrd_last_error.restype = C.c_char_p
L.cldrd_format_score_selfcheck.restype = C.c_int64
def P(a): return a.ctypes.data_as(C.c_void_p)
rng = np.random.default_rng(0)
d = tempfile.mkdtemp()
# writer: many shapes incl. k=1, k=2048, negative ids, threads
for nq, k, T in ((1, 1, 1), (3, 2048, 2), (500, 37, 8), (50, 1000, 3), (0, 5, 1)):
    D = (rng.standard_normal((nq, k)) * 10.0 ** rng.integers(-10, 10)).astype(np.float32)
    I = rng.integers(-2**63, 2**63 - 1, (nq, k), dtype=np.int64)
    q = rng.integers(-2**63, 2**63 - 1, nq, dtype=np.int64)
    n = C.c_int64()
    path = os.path.join(d, "w.tsv").encode()
    rc = L.cldrd_write_run_mt(path, P(q), P(D), P(I), C.c_int64(nq), k, 0, T, C.byref(n))
    assert rc == 0, L.cldrd_last_error()
    # reader on what the writer wrote
    nl, bad = C.c_int64(), C.c_int64()
    assert L.cldrd_read_run(path, None, None, C.c_int64(0), T, C.byref(nl), None) == 0
    assert nl.value == nq * k
    qq, pp = np.empty(nl.value, np.int64), np.empty(nl.value, np.int64)
    assert L.cldrd_read_run(path, P(qq), P(pp), C.c_int64(nl.value), T, C.byref(nl), C.byref(bad)) == 0, L.cldrd_last_error()
    assert np.array_equal(pp, I.reshape(-1)) and np.array_equal(qq, np.repeat(q, k))
# special scores
sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 3.4028235e38, -3.4028235e38, 2.0**-33, 2.0**53, 100.0, 0.1], dtype=np.float32)
buf = C.create_string_buffer(32)
for v in sp:
    n = L.cldrd_format_score(C.c_float(float(v)), buf)
    assert 0 < n <= 24
bad, fast = C.c_uint32(), C.c_int64()
assert L.cldrd_format_score_selfcheck(C.c_uint32(0), C.c_uint32(4099), C.c_int64(1 << 20), C.byref(bad), C.byref(fast)) == 0
# reader edge cases
for content in (b"", b"\n", b"1\t2", b"1\t2\n", b" \t \n", b"1\t2\t3\t4\t5", b"x" * 5000000, b"1\t2\n" * 100000 + b"3"):
    path = os.path.join(d, "r.tsv")
    open(path, "wb").write(content)
    nl, bad = C.c_int64(), C.c_int64()
    rc = L.cldrd_read_run(path.encode(), None, None, C.c_int64(0), 4, C.byref(nl), None)
    qq, pp = np.empty(max(nl.value, 1), np.int64), np.empty(max(nl.value, 1), np.int64)
    rc = L.cldrd_read_run(path.encode(), P(qq), P(pp), C.c_int64(nl.value), 4, C.byref(nl), C.byref(bad))
    print(len(content), "bytes ->", nl.value, "lines rc", rc, "bad", bad.value)
# ranged index writer + sync
w = C.c_void_p()
ip = os.path.join(d, "i.index").encode()
xb = rng.standard_normal((100, 8)).astype(np.float32); ids = np.arange(100, dtype=np.int64)
assert L.cldrd_index_writer_open_range(C.byref(w), ip, C.c_int64(100), 8, 1, 0, C.c_int64(0), C.c_int64(60), 1) == 0
assert L.cldrd_index_writer_append(w, P(xb), C.c_int64(60)) == 0
assert L.cldrd_index_writer_sync(w) == 0
w2 = C.c_void_p()
assert L.cldrd_index_writer_open_range(C.byref(w2), ip, C.c_int64(100), 8, 1, 0, C.c_int64(60), C.c_int64(40), 0) == 0
assert L.cldrd_index_writer_open_range(C.byref(C.c_void_p()), ip, C.c_int64(101), 8, 1, 0, C.c_int64(60), C.c_int64(40), 0) != 0
x2 = np.ascontiguousarray(xb[60:])
assert L.cldrd_index_writer_append(w2, P(x2), C.c_int64(40)) == 0
assert L.cldrd_index_writer_finish(w2, None) == 0
assert L.cldrd_index_writer_finish(w, P(ids)) == 0
print("asan host run ok")
