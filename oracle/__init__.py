"""CPU oracle for the SpeechMix speech-to-text hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker (or as the
CPU baseline being timed), never as the thing shipped.  The product path
(``speechmix_b200``) never imports this package and fails loudly when its
CUDA extension is missing.
"""
