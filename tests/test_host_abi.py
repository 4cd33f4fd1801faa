"""include/textgcn_host.h against libtextgcn_host.so: every declared symbol is exported, and a C program compiled
against the header reproduces the reference's known-answer test (textgcn/test/test_cfunc.py:81-99) and the golden
edge list of the same input."""
import ctypes
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "textgcn_host.h")
LIB = os.path.join(ROOT, "pytextgcn_b200", "lib", "libtextgcn_host.so")


def test_every_declared_symbol_is_exported():
    names = re.findall(r"\b(tgcn_ww_\w+)\s*\(", open(HEADER).read())
    assert set(names) == {"tgcn_ww_build", "tgcn_ww_fetch", "tgcn_ww_free", "tgcn_ww_counts_packed"}
    lib = ctypes.CDLL(LIB)
    for n in names:
        assert getattr(lib, n) is not None


def test_c_program_against_the_header_reproduces_the_reference_kat(tmp_path):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "graphbuilder_kat.npz"))
    X = gold["X"].astype(np.int32)                      # the 2 x 8 token matrix of test_cfunc.py:83-86 (-1 = padding)
    assert X.shape == (2, 8) and int(gold["n_vocab"]) == 6 and int(gold["window"]) == 3
    src = tmp_path / "kat.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "%s"
int main(void) {
  const int32_t X[16] = {%s};
  uint32_t c[21]; uint64_t nw = 0;
  if (tgcn_ww_counts_packed(X, 2, 8, 6, 3, c, &nw)) return 1;
  for (int i = 0; i < 21; ++i) printf("%%u ", c[i]);
  printf("\n");
  int64_t ne = 0;
  void* h = tgcn_ww_build(X, 2, 8, 6, 3, 1, &ne, &nw);
  if (!h) return 2;
  int32_t* coo = malloc(sizeof(int32_t) * 2 * ne); float* w = malloc(sizeof(float) * ne);
  if (tgcn_ww_fetch(h, coo, w)) return 3;
  for (int64_t e = 0; e < ne; ++e) printf("%%d %%d %%.8f\n", coo[2 * e], coo[2 * e + 1], w[e]);
  tgcn_ww_free(h);
  const int32_t bad[2] = {0, 9};
  return tgcn_ww_build(bad, 1, 2, 6, 2, 1, &ne, &nw) == NULL ? 0 : 4;     /* token id out of range -> NULL */
}''' % (HEADER, ", ".join(str(int(v)) for v in X.reshape(-1))))
    exe = tmp_path / "kat"
    subprocess.run(["gcc", "-o", str(exe), str(src), LIB, f"-Wl,-rpath,{os.path.dirname(LIB)}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert [int(v) for v in out[0].split()] == [int(v) for v in gold["c_ij"]]
    edges = [(int(a), int(b), float(w)) for a, b, w in (ln.split() for ln in out[1:])]
    assert [(a, b) for a, b, _ in edges] == [tuple(int(v) for v in r) for r in gold["coo"]]
    assert np.array_equal(np.array([w for _, _, w in edges], dtype=np.float32), gold["weights"].astype(np.float32))
