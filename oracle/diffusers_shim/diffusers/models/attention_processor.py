class LoRAAttnAddedKVProcessor:  # names only: the reference imports but never uses them on the path
    pass


class LoRAAttnProcessor:
    pass


class LoRAAttnProcessor2_0:
    pass


class LoRAXFormersAttnProcessor:
    pass


class SpatialNorm:
    pass
