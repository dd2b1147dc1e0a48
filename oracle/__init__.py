"""CPU oracle for the BPR hot path — TEST INFRASTRUCTURE ONLY.

Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it, and only as the checker or
the timed CPU baseline.  The product path (revisit-bpr_b200/) never imports this package and
fails loudly when librbpr.so is missing.

Parity status: the reference ships no tests or golden vectors, so the oracle
is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
tests/golden/make_golden.py (imports /root/reference unmodified, accelerate stubbed because
it is only a type annotation) and committed as tests/golden/*.npz.

  oracle.philox   numpy restatement of the counter-based negative sampler specification
                  (Philox4x32-10 + Lemire bounded draw + CSR rejection)  [bit-exact contract]
  oracle.ref_bpr  torch-CPU restatement of the reference op sequence: MF logits, BPR loss, L2,
                  autograd backward, dense torch.optim step, multinomial samplers, all-item
                  eval + seen masking + NDCG/Recall   [pinned against the reference]
  oracle.adaptive numpy restatement of the adaptive sampler; its deterministic part restates the
                  reference's formula and is pinned by tests/golden/adaptive.npz
  oracle.closed   numpy closed-form minibatch SGD step (independent of autograd; vectorised, so it
                  also checks BASELINE-size steps in seconds)
"""
