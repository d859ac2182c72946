"""Golden vectors for the N3 pieces: generated in the dev container by importing the UNMODIFIED reference
  * decoding.config.GenerationConfigCustom.update_from_string  (src/decoding/config.py:25-61)
  * utilities.generation_utils.save_nbests                     (src/utilities/generation_utils.py:16-52)
  * the eval-batch rescale of do_evaluate                       (src/utilities/general_utils.py:140-147, arithmetic only)
and writing tests/golden/n3_generation.json.  Run:  PYTHONPATH=/root/reference/src python tests/golden/make_golden_n3.py
"""
import json
import math
import os
import sys
import tempfile

import torch

sys.path.insert(0, "/root/reference/src")
from decoding.config import GenerationConfigCustom  # noqa: E402
from utilities.generation_utils import save_nbests  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ["ctc_weight", "ctc_margin", "lm_weight", "space_token_id", "eos_space_trick_weight", "apply_eos_space_trick", "num_beams",
          "max_length", "length_penalty", "early_stopping", "do_sample"]


def snapshot(c):
    return {k: getattr(c, k) for k in FIELDS}


def main():
    out = {"updates": [], "rescale": [], "nbests": None}
    base = dict(ctc_weight=0.3, ctc_margin=0, lm_weight=0.5, space_token_id=-1, eos_space_trick_weight=1.5, apply_eos_space_trick=False,
                num_beams=4, max_length=128, length_penalty=1.0, early_stopping=False, do_sample=False)
    cases = ["ctc_weight=0.5", "ctc_weight=0.3;num_beams=10", "apply_eos_space_trick=true;eos_space_trick_weight=2", "apply_eos_space_trick=Yes",
             "apply_eos_space_trick=0", "num_beams=8;max_length=64;length_penalty=0.5", "ctc_margin=3;space_token_id=17", "lm_weight=1",
             "early_stopping=n", "ctc_weight=1e-1", "no_such_key=1", "apply_eos_space_trick=maybe", "num_beams=2.5", "ctc_weight=abc",
             "num_beams=10;", "ctc_weight = 0.4", "num_beams=10,ctc_weight=0.2"]
    for s in cases:
        c = GenerationConfigCustom(**base)
        entry = {"base": base, "update": s}
        try:
            c.update_from_string(s)
            entry["result"] = snapshot(c)
        except Exception as e:  # noqa: BLE001
            entry["error"] = type(e).__name__
            entry["result"] = snapshot(c)  # keys applied before the failure stay applied
        out["updates"].append(entry)
    # general_utils.py:140-147
    for bs, orig, new in [(32, 4, 10), (32, 10, 4), (8, 1, 20), (7, 5, 5), (64, 4, 8), (3, 2, 9)]:
        out["rescale"].append({"eval_batch": bs, "beams_orig": orig, "beams_new": new,
                               "result": bs if new == orig else math.ceil(bs / (new / orig))})

    class Tok:
        pad_token_id = 3

        def decode(self, ids, skip_special_tokens=True):
            return " ".join(f"t{i}" for i in ids if not (skip_special_tokens and i in (0, 1, 2, 3)))

    nb = [torch.tensor([[0, 7, 8, 1, 3], [0, 7, 9, 1, 3], [0, 5, 1, 3, 3], [0, 5, 6, 6, 1]]), torch.tensor([[0, 11, 1], [0, 12, 1]])]
    sc = [torch.tensor([-0.5, -0.75, -1.25, -2.0]), torch.tensor([-0.125, -3.5])]
    lb = [torch.tensor([[7, 8, 1, -100], [5, 6, 1, -100]]), torch.tensor([[11, 1, -100]])]
    with tempfile.TemporaryDirectory() as d:
        save_nbests(os.path.join(d, "nb"), [t.clone() for t in nb], [t.clone() for t in sc], [t.clone() for t in lb], Tok(), group_size=2,
                    batch_size=2, outputs=None)
        files = {suf: open(os.path.join(d, "nb" + suf)).read() for suf in ("_scores.txt", "_hyps.txt", "_refs.txt")}
    out["nbests"] = {"nbests": [t.tolist() for t in nb], "scores": [t.tolist() for t in sc], "labels": [t.tolist() for t in lb], "group_size": 2,
                     "files": files}
    with open(os.path.join(HERE, "n3_generation.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote n3_generation.json:", len(out["updates"]), "update cases")


if __name__ == "__main__":
    main()
