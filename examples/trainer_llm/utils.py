import torch


def conv_str_to_dtype(s: str) -> torch.dtype:
    table = {"torch.float32": torch.float32, "torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16}
    if s not in table:
        raise ValueError(f"Unknown dtype {s}")
    return table[s]
