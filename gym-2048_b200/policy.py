"""The policy network the step path feeds in BASELINE config 5 (dense torch code, not part of
the hot path): the residual CNN trunk of `/root/reference/ppo_train.py:37-62`
(`ResNetExtractor`: conv3x3 16->filters + BN + ReLU, `residual_blocks` x `model.ResidualBlock`
(`/root/reference/model.py:10-26`), flatten) with the linear action and value heads SB3's
`ActorCriticCnnPolicy` attaches for `net_arch=[]` (ppo_train.py:125-133).  Random-init weights
(there is no checkpoint to load); cuDNN/cuBLAS do the arithmetic."""
import torch
import torch.nn as nn


class ResidualBlock(nn.Module):
    """Two conv3x3 + BatchNorm with a skip connection (model.py:10-26)."""

    def __init__(self, filters):
        super().__init__()
        self.conv1 = nn.Conv2d(filters, filters, kernel_size=3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(filters)
        self.conv2 = nn.Conv2d(filters, filters, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(filters)

    def forward(self, x):
        y = torch.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return torch.relu(y + x)


class ResNetActorCritic(nn.Module):
    """obs [B,16,4,4] -> (action logits [B,4], values [B])."""

    def __init__(self, filters=64, residual_blocks=4, board_layers=16, board_size=4, outputs=4):
        super().__init__()
        self.trunk = nn.Sequential(
            nn.Conv2d(board_layers, filters, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(filters),
            nn.ReLU(inplace=True),
            *[ResidualBlock(filters) for _ in range(residual_blocks)],
            nn.Flatten(),
        )
        features = filters * board_size * board_size
        self.action_net = nn.Linear(features, outputs)
        self.value_net = nn.Linear(features, 1)

    def forward(self, obs):
        f = self.trunk(obs.float() if obs.dtype not in (torch.float32, torch.bfloat16, torch.float16) else obs)
        return self.action_net(f), self.value_net(f).squeeze(-1)

    def flops_per_obs(self):
        """Multiply-add FLOPs (2 per MAC) of one forward pass for one observation."""
        total = 0
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                total += 2 * m.in_channels * m.out_channels * m.kernel_size[0] * m.kernel_size[1] * 16
            elif isinstance(m, nn.Linear):
                total += 2 * m.in_features * m.out_features
        return total
