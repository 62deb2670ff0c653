"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds the checkers for the demodulation hot path: the reference's own classes
compiled into `oracle/_ref/libfmref.so` (ref.py), a plain-C restatement
(`oracle/restate/`, restate.py) and the deterministic synthetic IQ generators (siggen.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
package; the product (airspy_fmradion_b200) never does.
"""
