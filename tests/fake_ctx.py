"""A CPU stand-in for twilight_b200.Context (TEST INFRASTRUCTURE): same rows_* / align_level surface, backed by the CPU
oracle, so the host-side drivers (progressive scheduling, multi-rank sharding and node migration) can be exercised
without a GPU. Never used by the product."""
import numpy as np

from tests import oracle_lib as ol, ref_msa
from twilight_b200 import api


class FakeContext:
    P = 6

    def __init__(self, cfg=None):
        self.cfg = cfg or ol.TalcoCfg()
        self.rows, self.w = {}, {}
        self._merged = []

    def rows_clear(self):
        self.rows.clear()
        self.w.clear()

    def rows_upload(self, ids, rows, weights):
        for i, r, w in zip(ids, rows, weights):
            self.rows[int(i)] = bytes(r)
            self.w[int(i)] = float(w)

    def rows_download(self, ids):
        return [self.rows[int(i)] for i in ids]

    def align_level(self, pairs, task=0, gappy=0.95, cache_threshold=1000):
        outs, self._merged = [], []
        for p in pairs:
            sides = []
            for sd in (p.ref, p.qry):
                sides.append(ref_msa.NodeState([self.rows[i] for i in sd.seq_ids], np.array([self.w[i] for i in sd.seq_ids], np.float32),
                                               sd.aln_len, sd.aln_num, sd.aln_weight, sd.msa_freq))
            rec = ref_msa.align_pair("n", self.cfg, sides[0], sides[1], gappy, task, None, cache_threshold)
            ids = list(p.ref.seq_ids) + list(p.qry.seq_ids)
            for i, r in zip(ids, rec.merged.rows):
                self.rows[i] = r
            self._merged.append(rec.merged.msa_freq)
            outs.append(api.LevelOut(rec.error, rec.aln_w, rec.tiles, rec.cells, False, False, rec.merged.msa_freq is not None,
                                     len(rec.profile[0]), len(rec.profile[1])))
        return outs

    def level_fetch(self, pair, what):
        return self._merged[pair]

    def level_phase_ms(self):
        return [0.0, 0.0, 0.0, 0.0]

    def launch_count(self):
        return 0
