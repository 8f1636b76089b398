"""Running mean with the interface of the reference's metrics.py:27-49 (``.val`` is the running average)."""


class AverageMeter:
    def __init__(self):
        self.reset()

    def reset(self):
        self.value = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.value = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    @property
    def val(self):
        return self.avg
