"""trainner_b200 -- B200-native (sm_100a) implementation of the ESRGAN training hot path of
victorca25/traiNNer: RRDBNet generator, VGG-style discriminator, VGG19 perceptual features and the
SRModel.optimize_parameters G/D step, behind the reference's nn.Module / arch-registry surface.

Importing the package does not need a GPU; running any network does (there is no CPU fallback --
the CPU oracle lives in oracle/ and is test infrastructure only).
"""
__version__ = "0.1.0"
