"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the keypoints2body fitting path.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline.  The product path (``keypoints2body_b200``) never imports this
package and fails loudly when its CUDA library is missing.

Contents
--------
``smplx_shim``      torch restatement of the third-party ``smplx`` forward
                    (smplx is unpinned in the reference's pyproject.toml:22,
                    ``>=0.1.28`` in environment.yaml:18, and is NOT installed
                    here) -- the only arithmetic that had to be restated.
``reference_port``  restatement of the reference's fitter / loss / prior on top
                    of torch autograd + torch.optim (the reference's own
                    optimisers), runnable anywhere torch is (incl. the GPU box,
                    where /root/reference does not exist).
``ref_loader``      imports the UNMODIFIED reference from /root/reference via
                    two stub modules (authoring container only); used by
                    ``tests/golden/make_goldens.py`` to pin ``reference_port``.

Parity status: PINNED -- ``reference_port`` is checked against golden vectors
produced by the unmodified reference code (``tests/golden/*.npz``, generator
``tests/golden/make_goldens.py``) in ``tests/test_oracle_vs_reference.py``.
The reference's own tests hold no numeric fixture for this path (SURVEY.md
section 4), so those goldens are the pin.
"""
