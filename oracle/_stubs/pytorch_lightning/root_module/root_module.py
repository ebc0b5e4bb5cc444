import torch.nn as nn


class LightningModule(nn.Module):
    use_ddp = False
