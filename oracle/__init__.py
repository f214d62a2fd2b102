"""CPU oracle for the ViT attention hot path of zhengyk19/vit-rpe-rope.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and there only as the checker
or as the CPU baseline being timed.  The product package
``vit_rpe_rope_b200`` never imports this package and has no CPU fallback.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so this oracle is pinned against *outputs of the
reference itself*: ``oracle/make_golden.py`` imports the unmodified reference
from ``/root/reference`` (behind a 3-symbol ``timm`` stub, ``oracle/_timm_stub.py``)
and writes the fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every oracle function against them (bit-exact for the integer and
coordinate tables and - on the same CPU/torch build - for the full model).

Contents
--------
* ``tables_np``      numpy restatement of the index / coordinate / frequency
                     tables (integer + fp32, bit-exact targets).
* ``attention_np``   numpy float64 restatement of the attention core, forward
                     and hand-derived backward (the formulas the CUDA kernels
                     implement).
* ``vit_torch``      torch-CPU fp32 restatement of the whole reference model
                     (the floating-point reference; also what
                     ``bench.py --impl reference`` times).
* ``vrr_oracle.c``   plain-C restatement of the tables and the attention core
                     (double precision), built by ``oracle/Makefile``.
"""
