"""visualize_sharding: print which rank holds which region of a 1-D/2-D DTensor."""
from __future__ import annotations


from ...layout import local_boxes

__all__ = ["visualize_sharding"]


def visualize_sharding(dtensor, header: str = "") -> str:
    mesh, spec = dtensor.device_mesh, dtensor._spec
    shape = tuple(dtensor.shape)
    if len(shape) not in (1, 2):
        raise RuntimeError("visualize_sharding supports 1-D and 2-D tensors")
    rows = [header] if header else []
    flat = mesh.mesh.flatten().tolist()
    import torch

    for idx, rank in enumerate(flat):
        coord = tuple(int(i) for i in torch.unravel_index(torch.tensor(idx), mesh.shape))
        boxes = local_boxes(shape, mesh, spec.placements, coord)
        desc = "; ".join(
            ",".join(f"{o}:{o + s}" for o, s in zip(off, sz)) for off, sz, _ in boxes
        ) or "(empty)"
        rows.append(f"rank {rank} @ {coord}: [{desc}]")
    out = "\n".join(rows)
    if mesh.get_rank() == flat[0]:
        print(out)
    return out
