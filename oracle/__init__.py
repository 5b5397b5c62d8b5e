"""CPU oracle for the AudioLCM latent->waveform decode path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as the thing that is shipped.  The product
package ``audiolcm_b200`` never imports this package.

Parity pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the *live reference
modules* instead: ``oracle/make_golden.py`` imports the unmodified reference
from ``/root/reference``, loads deterministic synthetic weights
(``audiolcm_b200/synth.py``) into it, and commits its outputs under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks this restatement against those vectors.
"""
