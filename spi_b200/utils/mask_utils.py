"""spi/utils/mask_utils.py:4-9."""
import torch

FACE_LABELS = (1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13)


def calculate_face_mask(mask):
    face_mask = torch.zeros_like(mask)
    for att in FACE_LABELS:
        face_mask += (mask == att)
    return face_mask
