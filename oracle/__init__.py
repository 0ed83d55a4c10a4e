"""CPU oracle for the speedplusbaseline CNN-training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``speedplusbaseline_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
(or as the CPU arm being timed), never as the product.

What it is: a restatement, in plain functional ``torch`` CPU ops, of the
reference's algorithm for the path named by BASELINE.json (KRN / SPN forward,
the Ghiasi style-augmentation net, the DANN gradient-reversal branch, the
clip-norm + AdamW step).  The reference (``/root/reference``, pure Python) puts
all arithmetic in a third-party dependency that is not vendored there:
``torch==1.8.0`` / ``torchvision==0.9.0`` (requirements.txt:4,7); this image has
torch 2.11.0 / torchvision 0.26.0, whose conv2d / batch_norm / AdamW math for
the call sites on the path is unchanged.  The oracle
therefore composes ``torch.nn.functional`` CPU ops exactly the way the
reference modules do, but from a flat ``{state_dict key: tensor}`` mapping, so
that it travels to the GPU box (``/root/reference`` does not exist there).

Parity pin: the reference ships no tests, golden vectors or KATs
("parity unpinned" upstream).  The pin used here is
outputs of the reference itself run in the build container:
``oracle/make_golden.py`` imports the unmodified reference modules, feeds them
the seeded synthetic weights/inputs of ``oracle/synth.py`` and stores the
results under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every
oracle function against those files.
"""
