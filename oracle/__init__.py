"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference's training-step algorithm.

Nothing in the product package (``frl-distributed-ml-scaffold_b200/``) imports this directory.
Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` do, and
there only as the checker / the reported CPU baseline — never as the thing shipped or timed as
the GPU path.
"""
