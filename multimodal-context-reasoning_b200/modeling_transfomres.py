"""Import-surface twin of the reference's ``modeling/modeling_transfomres.py`` (sic), from which the run
scripts import ``BertImgModel`` (run_PMR_ModCR.py:21).  The encoder classes live in ``modeling_bert``."""
from .modeling_bert import (BertImgModel, CaptionBertAttention, CaptionBertEncoder, CaptionBertLayer,  # noqa: F401
                            CaptionBertSelfAttention)
