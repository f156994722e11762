def warmup_linear_lambda(step: int, warmup_steps: int, t_total: int) -> float:
    """EasyNLPWarmupLinearSchedule.lr_lambda (easynlp/core/optimizers.py:191-204)."""
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    return max(0.0, float(t_total - step) / float(max(1.0, t_total - warmup_steps)))
