"""Pick the strongest available checker: the compiled reference (oracle/_ref/libfmref.so)
when it was built, else the plain-C restatement (oracle/restate)."""
import numpy as np

from oracle import ref


def have_ref():
    return ref.available()


def _restate():
    from oracle import restate
    return restate


def oracle_fm_run(iq, fs, blk, stereo=True, fs4=False, filter=0, deemphasis_us=50.0, pilot_shift=False,
                  mpf_stages=0, taps=(), prefer="ref"):
    """Run one channel block by block. Returns (audio, per_call_len) [+ tapdict if taps]."""
    if prefer == "ref" and have_ref():
        c = ref.RefChain("fm", fs, fs4=fs4, filter=filter, stereo=stereo, deemphasis_us=deemphasis_us,
                         pilot_shift=pilot_shift, mpf_stages=mpf_stages)
        audio, lens, td = c.run(iq, blk, taps=taps)
        st = c.stats()
        c.close()
    else:
        audio, lens, td, st = _restate().fm_run(iq, fs, blk, stereo=stereo, fs4=fs4, filter=filter,
                                                deemphasis_us=deemphasis_us, pilot_shift=pilot_shift,
                                                mpf_stages=mpf_stages, taps=taps)
    if taps:
        return audio, lens, td, st
    return audio, lens


def oracle_am_run(iq, fs, blk, filter=0, fs4=False, prefer="ref"):
    if prefer == "ref" and have_ref():
        c = ref.RefChain("am", fs, fs4=fs4, filter=filter)
        audio, lens, _ = c.run(iq, blk)
        c.close()
        return audio, lens
    audio, lens, _, _ = _restate().am_run(iq, fs, blk, filter=filter, fs4=fs4)
    return audio, lens


def oracle_nbfm_run(iq, fs, blk, filter=0, fs4=False, freq_dev=8000.0, prefer="ref"):
    """Returns (audio, per_call_len, stats)."""
    if prefer == "ref" and have_ref():
        c = ref.RefChain("nbfm", fs, fs4=fs4, filter=filter, freq_dev=freq_dev)
        audio, lens, _ = c.run(iq, blk)
        st = c.stats()
        c.close()
        return audio, lens, st
    audio, lens, _, st = _restate().nbfm_run(iq, fs, blk, filter=filter, fs4=fs4, freq_dev=freq_dev)
    return audio, lens, st
