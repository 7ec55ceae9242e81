"""CPU oracle for the CLIP-GP few-shot adapter hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``clip_gp_b200/`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may.  It is the checker, never the thing shipped.

What it restates (all citations are into ``/root/reference``):

* ``oracle.gp``       – ``trainers/gp_template_weigher.py`` (whole file) plus the
  semantics of the two un-vendored, un-pinned third-party packages that hold the
  GP arithmetic: ``gpytorch`` (+ ``linear_operator``; ``requirements.txt:8``) and
  ``entmax`` (``requirements.txt:13``).  Neither package is installable in the
  build container (no network), so their published algorithms are restated from
  the 1.11-1.14 gpytorch line / entmax 1.x.
* ``oracle.heads``    – the cosine-logit heads of ``trainers/adapter.py:200-252``
  and ``:387-476``, ``trainers/taskres.py:45-47,96-123``,
  ``trainers/clip_adapter.py:16-32,77-100``, and the Tip-Adapter cache affinity
  of ``trainers/tip_adapter.py:43-80,250-260``.
* ``oracle.metrics``  – ``utils/metrics.py:9-229`` (accuracy / ECE / AECE).

Pinning status (round 2)
-----------------------
Everything is pinned on vectors produced by EXECUTING THE REFERENCE'S OWN FILES in the build container
(``tests/golden/make_ref_golden.py``; the GPU box only sees the committed ``tests/golden/*.npz``):

* metrics: ``tests/golden/metrics_golden.npz`` = outputs of the reference's ``utils/metrics.py`` (imported by file path);
  the oracle matches bit-for-bit on counts and to 1e-6 on the float summaries.
* GP weighter: ``tests/golden/ref_gp.npz`` = the reference's ``trainers/gp_template_weigher.py`` imported UNMODIFIED on top of
  ``oracle/_shim`` (minimal dense-tensor stand-ins for the ~20 gpytorch / linear_operator / entmax routines it reaches; those
  libraries are un-vendored, un-pinned and not installable offline).  Setup (PCA, f0, median length-scale), first-call
  initialisation of q(u), mu / Sigma / w / prototypes / KL, every parameter gradient, the ``batch == K`` branch, eval-mode and
  no_grad variants, for rbf / matern / linear x four shapes (``tests/test_ref_golden.py``: 56 cases).
* heads, losses, trainers: ``tests/golden/ref_train.npz`` = the reference's own ``Trainer.train()`` of Adapter, TaskRes,
  CLIP-Adapter and Tip-Adapter(-F) (GP pre-training loops included), run end to end on a stand-in CLIP that returns cached
  features (``tests/golden/_fake_clip.py``); ``tests/test_ref_train_golden.py`` replays them with ``oracle/heads.py`` /
  ``oracle/train_step.py``: per-step losses, learning rates, final parameters, zero-shot and final accuracy / ECE / AECE / bins.
* What remains a restatement: the BODIES of the library routines inside ``oracle/_shim`` (each names the gpytorch /
  linear_operator / entmax routine it follows); there is no wheel of those libraries in the image to diff against.
* ``oracle.philox``: Random123 known-answer vectors for Philox4x32-10.
"""
