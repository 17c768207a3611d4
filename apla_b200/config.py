"""`AplaConfig`: the attribute-dict the reference passes around as `config` (an EasyDict built from
`model_params.adaptation.params`, params/finetune/dinov2/NABirds/vit_b/apla.yml:2-6).  It has to answer both
`hasattr(config, 'inds_path')` and `'inds_path' in config` (src/apla/apla_vit.py:12,77)."""


class AplaConfig(dict):
    def __init__(self, partial_size, inds_path=None, **extra):
        super().__init__(partial_size=partial_size, **extra)
        if inds_path is not None:
            self["inds_path"] = inds_path

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v
