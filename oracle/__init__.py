"""CPU oracle for the DDNM sampling hot path — TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a plain PyTorch-fp32 / numpy restatement of the
reference algorithm (wyhuai/DDNM @ 00b58ea).  It exists so that the CUDA engine
in ``ddnm_b200/`` can be checked on machines where ``/root/reference`` is absent
(the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline / ``--impl reference`` legs may import it; the product path never
does (``ddnm_b200`` raises if its CUDA library is missing).

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the unmodified reference
from ``/root/reference`` in the build container and (a) asserts the restatement
agrees with it, (b) writes the fixtures under ``tests/golden/`` which
``tests/test_oracle_golden.py`` re-checks on every run.
"""
