from .gentime_watermark import (GentimeWatermark, SeedStrategy, SplitStrategy,  # noqa: F401
                                create_watermarker_from_string)
