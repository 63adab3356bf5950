"""TEST-ONLY stand-in for Mom2Accumulator built on the oracle, so the host-side sharding / reduce /
cache logic of layer_stats can be exercised with gloo on a GPU-less machine.  Never used by the product."""
import numpy as np
import torch

from oracle import emcid_oracle as orc


class OracleAccumulator:
    def __init__(self, d, h, act):
        self.d, self.h, self.act = d, h, act
        self.st = orc.SecondMomentOracle()

    def set_weights(self, W1, b1):
        self.W1, self.b1 = W1.detach().cpu().numpy(), b1.detach().cpu().numpy()

    def add(self, X, valid=None):
        X = X.detach().cpu().numpy().reshape(-1, self.h)
        if valid is not None:
            X = X[valid.detach().cpu().numpy().reshape(-1) != 0]
        self.st.add(orc.fc2_input(X, self.W1, self.b1, self.act))

    def finalize(self):
        m = self.st.mom2 if self.st.mom2 is not None else np.zeros((self.d, self.d), np.float32)
        return torch.from_numpy(m.copy()), torch.tensor(self.st.count, dtype=torch.int64)

    # accumulator state of a resumable pass (Mom2Accumulator.export_state / import_state): packed lower triangle + count
    def export_state(self):
        m = self.st.mom2 if self.st.mom2 is not None else np.zeros((self.d, self.d), np.float32)
        return torch.from_numpy(m[np.tril_indices(self.d)].astype(np.float64)), torch.tensor(self.st.count, dtype=torch.int64)

    def import_state(self, packed, count):
        m = np.zeros((self.d, self.d), np.float32)
        m[np.tril_indices(self.d)] = packed.numpy().astype(np.float32)
        self.st.mom2 = m + np.tril(m, -1).T
        self.st.count = int(count)

    def close(self):
        pass
