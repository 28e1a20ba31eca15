"""Searches (data seed, stop bias) cases for the stop-index bit-exactness test (SURVEY.md 7.3-1 / VERDICT r1 item 4b):
B = 8 utterances that ALL stop before max_len, at several different frames, with an oracle stop margin (minimum |stop logit| over
the valid frames of the whole batch) at least 2x the bf16 path's logit error (~0.005; the GPU test asserts the ratio live).  Writes tests/golden/stop_cases.json with
the oracle's lengths, so the GPU test needs no search.  Run on CPU:  python tests/golden/search_stop_cases.py"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import synthetic  # noqa: E402

B, S, MAX_LEN, WANT, MIN_MARGIN = 8, 20, 72, 10, 0.016


def main():
    torch.set_num_threads(4)
    cases = []
    for bias in (-0.45, -0.5, -0.4, -0.55):
        m = synthetic.make_model(stop_bias=bias)
        for trial in range(400):
            ds, seed = 300 + trial, 1 + trial          # the dropout seed shapes the trajectory more than the phonemes do
            ph, pl, _, _ = synthetic.make_inputs(B, S, 8, ds, ragged=True)
            ma, lens, st = m.inference(ph, pl, max_len=MAX_LEN, seed=seed)
            T = st.shape[1]
            valid = torch.arange(T)[None, :] < lens[:, None]
            margin = float(st.abs()[valid].min())
            distinct = len(set(lens.tolist()))
            if int(lens.max()) < MAX_LEN and distinct >= 4 and margin >= MIN_MARGIN and int(lens.min()) >= 2:
                cases.append(dict(data_seed=ds, seed=seed, stop_bias=bias, lens=lens.tolist(), margin=margin))
                print(cases[-1], flush=True)
                if len(cases) >= WANT:
                    break
        if len(cases) >= WANT:
            break
    json.dump(dict(B=B, S=S, max_len=MAX_LEN, cases=cases), open(os.path.join(HERE, "stop_cases.json"), "w"), indent=1)
    print("found", len(cases))


if __name__ == "__main__":
    main()
