"""TensorBoard scalar logger with the reference's interface (logger.py:3-15 there)."""


class Logger(object):
    def __init__(self, log_dir):
        from torch.utils.tensorboard import SummaryWriter
        self.writer = SummaryWriter(log_dir)

    def scalar_summary(self, tag, value, step):
        self.writer.add_scalar(tag=tag, scalar_value=value, global_step=step, new_style=True)

    def histo_summary(self, tag, values, step, bins=1000):
        self.writer.add_histogram(tag, values, global_step=step, bins=bins)
