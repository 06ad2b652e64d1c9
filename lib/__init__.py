"""Drop-in package with the reference's module paths (`lib.models.SHAS`, `lib.evaluate.infer`, ...)
so that saved Hydra configs (`_target_: lib.models.SHAS`) and user scripts keep working; the
compute behind it is wav2vecsegmenter_b200 (sm_100a CUDA through libw2vseg.so)."""
