"""CPU oracle for the RAFT forward/update hot path of daigo0927/tf-raft.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`tf_raft_b200/`) may
import from here; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do, and there only as
the checker or as the timed CPU baseline -- never as the thing shipped.

PARITY UNPINNED (model level): the reference is pure Python on TensorFlow 2.3 +
tensorflow-addons 0.11, neither of which is installed here nor installable (no
network), so the reference itself cannot be executed to generate vectors, and
its own tests hold no golden values for correlation, lookup, update block or
final flow (shape checks only, tests/test_model.py:44-77).  What *is* pinned:

  * tests/layers/test_corr.py:15-27   sampler == standard bilinear on in-range,
                                      non-integer coords (atol 1e-5)
  * tests/test_model.py:14-41         extract_patches / depth_to_space orders
  * tests/losses/test_losses.py:27-67 sequence_loss / end_point_error answers

Those three are reproduced against this oracle in tests/test_oracle_pins.py.
Everything else rests on two independent restatements (a literal NumPy one,
`oracle.tf_ops` + `oracle.corr_np`, and a PyTorch-CPU functional one,
`oracle.raft_torch`) agreeing with each other, each function citing the
reference file:line it follows.
"""
