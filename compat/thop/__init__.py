"""Import-time stub for ``thop`` (models_bid_pointconv.py:680 imports it at module scope; it is only
used inside ``if __name__ == '__main__'``).  Not a MAC counter."""


def profile(model, inputs=(), **kwargs):
    params = float(sum(p.numel() for p in model.parameters()))
    return 0.0, params


def clever_format(nums, fmt="%.2f"):
    return [fmt % float(n) for n in nums]
