"""CPU oracle for the MaxViT / MetNet3 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / reference
legs of ``bench.py`` may import it, and only as the checker (or as the timed
CPU baseline), never as the thing shipped.  The product path
(``vit-grid-model_b200``) never imports this package and has no CPU fallback.

The functions here are a functional (state-dict in, tensor out) restatement of
the reference modules in plain PyTorch fp32:

* ``maxvit_oracle``  – /root/reference/src/maxvit.py
* ``metnet3_oracle`` – /root/reference/src/metnet3.py:86-430
* ``focal_r_oracle`` – not in the reference (README.md:16 only): parity unpinned

Pinning: ``tests/golden/make_golden.py`` imports the *real* reference modules
(in the build container, where /root/reference exists), loads the synthetic
state dict of ``oracle.synth`` with ``strict=True`` and stores the reference's
outputs as small fixtures under ``tests/golden/``; ``tests/test_oracle.py``
checks this restatement against those fixtures.
"""
