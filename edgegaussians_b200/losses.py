"""Mirror of /root/reference/edgegaussians/models/losses.py:5-11 (device-resident masks/weights)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MaskedL1Loss(nn.Module):
    def forward(self, input, target, mask):
        return F.l1_loss(input[mask], target[mask])


class WeightedL1Loss(nn.Module):
    def forward(self, input, target, weights):
        return torch.mean(weights.to(input.device) * torch.abs(input - target))
