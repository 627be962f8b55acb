"""TEST INFRASTRUCTURE ONLY.  Stand-in for the three xformers operators the reference's Chameleon transformer imports
(deps/chameleon/inference/transformer.py:12-15), written from their documented semantics in plain torch so that the
reference's OWN `Transformer` module can be imported and run on the CPU of the build container (xformers itself is not
installed and has no CPU kernels for these ops).  Used only by oracle/gen_golden_chameleon_transformer.py."""
