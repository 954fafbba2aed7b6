"""CPU oracle of the reference's classification training step.

TEST INFRASTRUCTURE ONLY.  These are plain-PyTorch fp32 restatements of the reference models' arithmetic
(KKKSQJ/DeepLearning, classification/*), written as stateless functions over a ``state_dict`` with the reference's key
names.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this package; the
product path (``deeplearning_b200``) never does and fails loudly when its CUDA extension is missing.

Pinning: the reference publishes no golden vectors for this path, so the restatements are pinned
against the reference itself, imported from /root/reference in the build container by ``tests/golden/make_golden.py``,
which (a) asserts bit-identical outputs between each oracle function and the reference module on the same weights and
inputs and (b) writes the small fixtures in ``tests/golden/`` that the CPU test-suite replays without the reference.
"""
