"""profiles/k1_imin.json (the I_min of bench.py's roofline.achieved_algorithmic) is what tests/imin_audit.py derives
from the oracle's draw counts; regenerate with  python -m tests.test_instruction_budget  after changing the audit."""
import json
import os

from tests import imin_audit

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "k1_imin.json")


def test_committed_budget_matches_the_audit():
    committed = json.load(open(PATH))
    assert committed["cost_per_item_sass_instructions"] == imin_audit.COST
    for name in ("cfg1", "cfg2"):
        got = imin_audit.audit(name, n_sims=120, seed=5000)      # other simulations than the committed sample
        ref = committed["per_config"][name]
        for k in ("visits", "btrs", "inv", "poisson", "I_min_thread_inst_per_reaction"):
            assert abs(got[k] - ref[k]) <= 0.06 * ref[k], (name, k, got[k], ref[k])
        assert 1.1 < ref["trials_per_btrs_draw"] < 1.35


if __name__ == "__main__":
    out = {"what": imin_audit.__doc__.split("\n\n")[0], "cost_per_item_sass_instructions": imin_audit.COST,
           "per_config": {n: imin_audit.audit(n) for n in imin_audit.CONFIGS}}
    json.dump(out, open(PATH, "w"), indent=1)
    print(json.dumps(out["per_config"], indent=1))
