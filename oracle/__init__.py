"""CPU oracle for the ADAPTed boundary-detection hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import anything from here.  The product (``adapted_b200``) never does: it fails loudly when
its CUDA library is missing instead of falling back to this code.

Contents
--------
``detect_ref.py``   numpy/scipy/torch-CPU restatement of the reference hot path
                    (adapted/detect/*.py, adapted/partition/signal_partitions.py).
``llr_gains.c``     plain-C restatement of adapted/detect/_c_llr.pyx:22-236 (gcc, no FMA).
``bn_restate.py``   float32 restatement of bottleneck.move_mean / move_var (third-party, absent from
                    the image; PARITY UNPINNED for this dependency).
``build_ref.py``    builds ``oracle/_ref`` (the reference's own Cython kernel compiled from
                    /root/reference) and a patched scratch copy of the python reference under /tmp.
``make_golden.py``  runs the *real* reference on seeded synthetic minibatches and writes
                    ``tests/golden/*.npz``.

Pinning status: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the reference itself executed in the build container
(``tests/golden`` + ``tests/test_oracle_vs_reference.py``).  Third-party arithmetic that is absent
from the image (bottleneck, pod5 calibration) is restated from its published algorithm and is
"parity unpinned".
"""
