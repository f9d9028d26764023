"""CPU oracle for the blurred-GAN hot path (blur + SWD).

TEST INFRASTRUCTURE ONLY.  Nothing under ``blurred_gan_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the CPU baseline -- never as the product path.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the
reference's own code run in the build container (``tests/golden/make_golden.py``
exec's ``pytorch/gaussianblur.py:1-72`` and imports ``sliced_wasserstein.py``
unmodified) and the resulting fixtures are committed under ``tests/golden/``.
The TF blur (``gaussian_blur.py``) cannot run here (no TensorFlow); its
restatement is pinned only by self-consistency (impulse response, explicit-tap
agreement with the torch oracle) and says so in ``blur_oracle.py``.
"""
