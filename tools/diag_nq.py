import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]
import torch
from cldrd import dist as CD
rng = np.random.Generator(np.random.PCG64(500))
xb = torch.from_numpy(rng.standard_normal((1_200_000, 64), dtype=np.float32)).cuda()
for nq in (4700, 8192, 8064, 6980):
    xq = torch.from_numpy(rng.standard_normal((nq, 64), dtype=np.float32)).cuda()
    s = CD.ShardedSearcher.from_rows(xb, 0, xb.shape[0], scan="f16")
    for k in (100, 1000):
        D, I = s.local.search_device(xq, k, translate_ids=False)
        st = s.shard.stats()
        print(nq, k, st, "surv/q", st["survivors"] / nq, flush=True)
    # by hand: sample -> seed -> seeded search, look at how many rows came back per query
    topj = s.local.sample_device(xq, 100)
    print(" topj[0,:4]", topj[0, :4].tolist(), "topj min of col31", float(topj[:, 31].min()), "max", float(topj[:, 31].max()),
          "n -inf", int(torch.isinf(topj[:, 31]).sum()))
    s.shard.close()
