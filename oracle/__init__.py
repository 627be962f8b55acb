"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

A restatement of the reference's algorithm for the hot path, used as the parity checker by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  Nothing under
``wmar_b200/`` imports this package: the product path is CUDA only and fails loudly without its extension.

Layout
  wm_oracle.c   integer work: MT19937 greenlist split, context seed, n-gram detector   (built by ``make``)
  wm.py         ctypes binding for wm_oracle.c + p-value (scipy.special.betainc, the reference's own call)
  sampling.py   one sampling step: +delta on green, /T, top-k, top-p, softmax, argmax(p/q)
  gpt.py        Taming minGPT decode step / sampling loop (fp32)
  rar.py        RAR decode step / generate loop with CFG (fp32)
  vqgan.py      Taming VQGAN and MaskGIT-VQGAN encode / decode / VQ argmin (fp32)
  gen_golden.py imports the real reference from /root/reference and writes tests/golden/*.npz

Parity pinning: the reference ships no tests or golden vectors for this path, so every
restatement here is pinned against outputs of the *imported reference code itself* on seeded inputs, stored
under tests/golden/ together with the generating script (oracle/gen_golden.py).
"""
