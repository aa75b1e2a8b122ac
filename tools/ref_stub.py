"""Uniform stub evaluator with the interface the reference search calls
(mcts.py:203-210, network.py:87-105): value 0, uniform moves_logprob over the
legal moves (padding masked with -99 like network.py:150).  Module-level so
that the reference's spawn workers can unpickle it."""
import torch


class UniformStub:
    device = torch.device('cpu')

    def eval(self):
        return self

    def run(self, batch):
        legal = batch['legal_moves']
        logit = torch.where(legal > 0, torch.zeros(legal.shape), torch.full(legal.shape, -99.0))
        return {'value': torch.zeros(len(legal)),
                'moves_logprob': torch.log_softmax(logit, dim=1)}
