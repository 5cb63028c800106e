"""Host-side helpers the hot path needs (NestedTensor batching, box geometry, distributed probes).
They mirror the call signatures of the reference's util.misc / util.box_ops for the functions the
models/dino package imports; everything else of the reference's util/ stays the reference's."""
