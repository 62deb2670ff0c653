import numpy as np, sys
sys.path.insert(0,'.')
from oracle import siggen
from tests.oracle_select import oracle_fm_run
from airspy_fmradion_b200 import FmDecoder
fs, blk, nblk = 384000.0, 2048, 40
iq = siggen.fm_stereo_iq(fs, blk * nblk, 1)[None, :]
for kw in (dict(filter=1), dict(filter=2), dict(pilot_shift=True), dict()):
    for percall in (nblk, 1):
        dec = FmDecoder(fmfilter=kw.get("filter", 0), stereo=True, pilot_shift=kw.get("pilot_shift", False), input_rate=fs, n_channels=1, max_samples_per_call=blk * nblk)
        outs=[]
        for o in range(0,nblk,percall):
            a,l = dec.process_blocks(iq[:, o*blk:(o+percall)*blk], [blk]*percall); outs.append(a)
        audio=np.concatenate(outs,axis=1)
        ref_audio, ref_lens = oracle_fm_run(iq[0], fs, blk, stereo=True, **kw)
        d = np.abs(audio[0]-ref_audio)
        print(kw, percall, "max %.3e at %d of %d" % (d.max(), d.argmax(), len(d)), "n>1e-6:", int((d>1e-6).sum()))
