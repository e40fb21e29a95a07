"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the IslaMcts simulate loop.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import it, and only as the checker / the CPU
baseline -- never as the thing measured or shipped.  The product package
``islamcts_b200`` must not import this package (tests/test_abi.py::test_product_never_imports_oracle checks).

Contents
--------
pyenvs.py      float64 Python restatement of the four Gymnasium envs
               (SURVEY.md section 8c -- env parity is UNPINNED by the reference:
               gym/gymnasium are third-party, unpinned, and absent here).
ref_loader.py  imports the reference's own agents from /root/reference with
               stub graphviz/gym/gymnasium modules (this container only).
trace.py       decision-trace recorder around the reference's random calls.
gen_golden.py  runs the reference and freezes traces + trees in tests/golden/.
py_port.py     pure-Python port of the simulate loop (CPU baseline, "port": the fallback).
ref_bench.py   stages the reference into baseline/_ref and times its own classes (CPU baseline, "reference").
c/             plain-C restatement (envs, Philox, search) -> liboracle.so.
coracle.py     ctypes binding of liboracle.so.
"""
