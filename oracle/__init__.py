"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the dense-correspondence matching path.

Nothing under oracle/ is imported by the product (midvision-probe_b200/).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it, and only as the
checker or the timed CPU baseline.

Parity status: **parity unpinned** in the sense of SURVEY.md section 8(c) -- the reference ships no
tests, golden vectors or known-answer files for this path, and its k-NN backend (faiss-gpu 1.8.0,
README.md:60; call sites evals/utils/correspondence.py:11, :20-22) is a third-party dependency that is
absent from /root/reference and from this image.  What pins the oracle instead:
  * oracle/reference_loader.py imports the reference's own, unmodified evals/utils/correspondence.py
    from /root/reference with an exact brute-force stand-in for the four faiss symbols it touches;
  * oracle/make_golden.py runs that module on seeded inputs and commits the outputs under
    tests/golden/ (the reference tree does not travel to the GPU box);
  * oracle/restated.py -- the restatement used everywhere else -- is checked against both.
"""
