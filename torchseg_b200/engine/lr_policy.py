"""LR policies (/root/reference/furnace/engine/lr_policy.py)."""


class BaseLR(object):
    def get_lr(self, cur_iter):
        raise NotImplementedError


class PolyLR(BaseLR):
    """lr_policy.py:18-26: lr0 * (1 - it/total)^power"""

    def __init__(self, start_lr, lr_power, total_iters):
        self.start_lr = start_lr
        self.lr_power = lr_power
        self.total_iters = total_iters + 0.0

    def get_lr(self, cur_iter):
        return self.start_lr * ((1 - float(cur_iter) / self.total_iters) ** self.lr_power)


class MultiStageLR(BaseLR):
    """lr_policy.py:29-38: [(iters, lr), ...]"""

    def __init__(self, lr_stages):
        assert type(lr_stages) in [list, tuple] and len(lr_stages[0]) == 2, \
            'lr_stages must be list or tuple, with [iters, lr] format'
        self._lr_stages = lr_stages

    def get_lr(self, epoch):
        for it_lr in self._lr_stages:
            if epoch < it_lr[0]:
                return it_lr[1]


class LinearIncreaseLR(BaseLR):
    """lr_policy.py:41-49"""

    def __init__(self, start_lr, end_lr, warm_iters):
        self._start_lr = start_lr
        self._delta_lr = (end_lr - start_lr) / warm_iters

    def get_lr(self, cur_epoch):
        return self._start_lr + cur_epoch * self._delta_lr
