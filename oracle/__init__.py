"""oracle/ -- TEST INFRASTRUCTURE ONLY. Never imported by the product path.

CPU restatement (numpy) of the arithmetic the reference reaches through four
``cv2`` calls, plus a thin harness that calls the live ``cv2`` wheel with the
reference's exact call-site arguments.

The reference (spirinis/HackathonOpticalFlow) holds no arithmetic of its own on
this path: everything runs inside the un-vendored third-party wheel
``opencv-python~=4.10.0.84`` (reference ``requirements.txt:1``).  This image
carries ``opencv-python-headless 4.13.0.92``; that wheel is the authority the
restatement is pinned against (``tests/test_oracle_*.py`` and the committed
fixtures under ``tests/golden/`` produced by ``tests/golden/make_golden.py``).

Parity pin status: the reference owns no tests, golden vectors or fixtures for
this path, so the pin is "outputs of the reference's own
cv2 calls run here" -- live cv2 on identical inputs -- not a reference-owned
vector set.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package.
"""
