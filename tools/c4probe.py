import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from haplohyped_varawareml_b200 import capi
for V, S, mix in ((12500, 200000, 1), (12500, 200000, 0), (1100000, 2504, 0), (1100000, 2504, 1)):
    sp = capi.synth_spec(V, S, seed=1000, mix=mix)
    T = int(capi.lib().hb_synth_body_bytes(sp))
    text = torch.empty(T + 256, dtype=torch.uint8, device="cuda"); text[T:].zero_()
    capi.check(capi.lib().hb_synth_device(sp, text.data_ptr(), T, 0, None))
    p = capi.Parse.from_device(text.data_ptr(), T, S, region="chr22")
    p.rerun(); p.rerun()
    i = p.info
    print("V", V, "S", S, "mix", mix, "used", i.tokenizer_used, "n_rec", i.n_records, "ms tok %.3f sites %.3f decode %.3f" % (i.ms_tokenize, i.ms_sites, i.ms_decode), "GB %.2f" % (T/1e9), flush=True)
    p.close(); del text
