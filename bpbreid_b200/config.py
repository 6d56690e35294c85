"""cfg.model.bpbreid node with the reference's defaults (torchreid/scripts/default_config.py:45-58), as a plain
attribute-dict so that the benchmark, the smoke test and user scripts can build the model without yacs.  A real
yacs CfgNode from the reference works just as well: the model only reads attributes."""


class Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def make_cfg(backbone="hrnet32", parts_num=5, **over):
    b = Cfg(pooling="gwap", normalization="identity", mask_filtering_training=False, mask_filtering_testing=True,
            last_stride=1, dim_reduce="after_pooling", dim_reduce_output=512, backbone=backbone,
            learnable_attention_enabled=True, test_embeddings=["bn_foreg", "parts"],
            test_use_target_segmentation="none", training_binary_visibility_score=True,
            testing_binary_visibility_score=True, shared_parts_id_classifier=False, hrnet_pretrained_path="",
            masks=Cfg(parts_num=parts_num))
    b.update(over)
    return Cfg(model=Cfg(bpbreid=b))
