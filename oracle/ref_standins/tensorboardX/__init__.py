from torch.utils.tensorboard import SummaryWriter  # noqa: F401
