"""CPU oracle for the SVS-UNet separation hot path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement of the reference algorithm for the path
STFT -> UNet mask -> mask x mixture -> iSTFT.  It is the *checker*: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``svs-unet-pytorch_b200/`` imports it, and the product path raises when the CUDA
library is missing instead of falling back to this code.

Parity status
-------------
* UNet half (``unet_oracle``): PINNED against the reference's own ``model.py``
  (imported in the build container from /root/reference with a stub for the
  unused ``import auraloss``); golden vectors + generator script live in
  ``tests/golden/``.
* Spectral half (``stft_oracle``): the arithmetic lives in the third-party
  dependency ``librosa==0.10.1`` (reference ``uv.lock:713-714``; -> scipy.fft
  pocketfft + two numba loops) which is absent from /root/reference and not
  installable here.  The reference has no tests / golden vectors for it
  (SURVEY.md section 4), so this half is "parity unpinned" at the librosa
  boundary: it restates librosa 0.10.1's published algorithm at the reference's
  call sites (data.py:79-85,100-105,151-164) and is cross-checked against the
  independent ``torch.stft`` / ``torch.istft`` implementations in float64.
"""
