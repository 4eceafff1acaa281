"""Model factory with the reference's signature and config keys (models/build_model.py:9-48).

    model, args = build_model(args_dict)                       # training: vars(config.args)
    model, args = build_model(None, load_config_dict=cfg)      # generate.py:340-341 / train.py:159

Required keys: vocab_size, n_layer, n_head, d_model, d_inner, dropout, d_condition, conditioning;
`regression` defaults to False (:26-27); `overwrite_dropout` is read only when load_config_dict is
given (:43-44).  max_seq = 2048 and pad_token = 0 are fixed like the reference (:22-23).
"""
from __future__ import annotations

from .transformer import MAX_SEQ, MusicTransformer, set_dropout

CONDITIONINGS = ("none", "discrete_token", "continuous_token", "continuous_concat")


def build_model(args, load_config_dict=None):
    if load_config_dict is not None:
        args = load_config_dict
    if args is None:
        raise ValueError("build_model needs an args dict or a load_config_dict")
    if "regression" not in args:
        args["regression"] = False
    conditioning = args["conditioning"]
    if args["regression"]:
        # models/build_model.py:29-32: MusicRegression with output_size 2 (it asserts d_condition <= 0)
        from .regression import MusicRegression
        model = MusicRegression(
            embedding_dim=args["d_model"], d_inner=args["d_inner"], d_condition=args["d_condition"],
            vocab_size=args["vocab_size"], num_layer=args["n_layer"], num_head=args["n_head"], max_seq=MAX_SEQ,
            dropout=args["dropout"], pad_token=0, output_size=2)
    else:
        if conditioning not in CONDITIONINGS:
            raise ValueError(f"unknown conditioning {conditioning!r}; expected one of {CONDITIONINGS}")
        model = MusicTransformer(
            embedding_dim=args["d_model"], d_inner=args["d_inner"], d_condition=args["d_condition"],
            vocab_size=args["vocab_size"], num_layer=args["n_layer"], num_head=args["n_head"], max_seq=MAX_SEQ,
            dropout=args["dropout"], pad_token=0, continuous_token=(conditioning == "continuous_token"))
    if load_config_dict is not None and args["overwrite_dropout"]:
        set_dropout(model, args["dropout"])
        print(f"Dropout rate changed to {args['dropout']}")
    return model, args
