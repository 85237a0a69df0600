"""Deterministic stand-in for the two nets, a function of the position only -- TEST INFRASTRUCTURE.
tests/golden/make_golden_mcts.py drives the REFERENCE search with `fake_nets`, tests/test_gpu_mcts.py drives the batched
tree of bokego_b200.mcts with the same function (through a subclass that overrides the net call), so tree statistics can be
compared count for count with the nets taken out of the picture."""
import numpy as np
import torch


def fake_nets(boards, turn):
    """Deterministic stand-in for the two nets, a function of the position only (test hook: tests/golden/make_golden_mcts.py
    drives the REFERENCE search with the same function, so the tree statistics can be compared exactly).
    boards int8 [n,81], turn int16 [n] -> (probs float32 [n,81], value float32 [n])"""
    import zlib
    probs = np.zeros((len(boards), 81), np.float32)
    val = np.zeros(len(boards), np.float32)
    for i in range(len(boards)):
        rng = np.random.RandomState(zlib.crc32(np.ascontiguousarray(boards[i], np.int8).tobytes() + bytes([int(turn[i]) & 0xFF])))
        lg = (rng.standard_normal(81) * 1.5).astype(np.float32)
        e = np.exp(lg - lg.max())
        probs[i] = (e / e.sum()).astype(np.float32)
        val[i] = np.float32(np.tanh(rng.standard_normal()))
    return probs, val



def fake_net_tree(mcts_module):
    """bokego_b200.mcts.MCTS whose net call is replaced by fake_nets (the encoder still supplies legal moves and the
    liberty cache)"""
    from bokego_b200.batched import features_batch

    class FakeNetMCTS(mcts_module.MCTS):
        def _bind_nets(self, policy_net, value_net):
            self.policy = self.value = None
            self.has_value = value_net is not False         # value_net=False: "a policy net but no value net"

        def _net_outputs(self, sub):
            out = features_batch(sub, want=("legal", "libs"))
            probs, val = fake_nets(sub.boards.cpu().numpy(), sub.turn.cpu().numpy())
            # what the reference's search sees: Categorical renormalises the probabilities in float32 (nnet.py:274)
            probs = torch.distributions.Categorical(probs=torch.from_numpy(np.asarray(probs, np.float32))).probs.numpy()
            return probs, np.asarray(val, np.float64), out["legal"].cpu().numpy(), out["libs"]

    return FakeNetMCTS


def fake_net_sim_tree(mcts_module, seed):
    """FakeNetMCTS for --simulate mode: playouts from the leaves run on the device with bk_playout_step, the fake policy and
    the draws of tests/golden/make_golden_mcts_sim.py -- the kernels' counter-based Exp(1) stream keyed (seed, playout number,
    turn, try 0), the same vector for every redraw of a move"""
    from bokego_b200 import batched as bk

    class FakeNetSimMCTS(fake_net_tree(mcts_module)):
        def _playout_results(self, leaves, first_id):
            out = np.zeros(len(leaves), np.float64)
            for j, leaf in enumerate(leaves):
                idx = torch.as_tensor([leaf], dtype=torch.long, device=self.device)
                p = self.pool
                pos = bk.Positions(p.boards[idx].contiguous(), p.ko[idx].contiguous(), p.last[idx].contiguous(),
                                   p.turn[idx].contiguous(), p.libs[idx].contiguous())
                pos.done = ((pos.turn > 80) | (pos.last == -1)).to(torch.uint8)
                while not bool(pos.done[0]):
                    turn = int(pos.turn[0])
                    probs, _ = fake_nets(pos.boards.cpu().numpy(), pos.turn.cpu().numpy())
                    probs = torch.distributions.Categorical(probs=torch.from_numpy(probs)).probs.to(self.device).contiguous()
                    q = bk.exp_draws(seed, first_id + j, turn, 0, 1, self.device)
                    mv = bk.playout_step(pos, probs, bk.MODE_MCTS, 80, q_inj=q.reshape(1, 1, 81).expand(1, 82, 81).contiguous())
                    assert int(mv[0]) >= -1
                _, reward = bk.score_batch(pos.boards)
                out[j] = float(reward[0])
            return out

    return FakeNetSimMCTS
