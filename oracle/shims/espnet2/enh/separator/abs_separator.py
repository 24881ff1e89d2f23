"""Stand-in for ``espnet2.enh.separator.abs_separator.AbsSeparator`` (a bare ``nn.Module`` ABC). TEST ONLY."""
import torch.nn as nn


class AbsSeparator(nn.Module):
    pass
