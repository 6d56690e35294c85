"""Dictionary keys shared by the model, the loss and the engine (same strings as the reference,
torchreid/utils/constants.py:1-19)."""
GLOBAL = "globl"
FOREGROUND = "foreg"
BACKGROUND = "backg"
CONCAT_PARTS = "conct"
PARTS = "parts"
BN_GLOBAL = "bn_globl"
BN_FOREGROUND = "bn_foreg"
BN_BACKGROUND = "bn_backg"
BN_CONCAT_PARTS = "bn_conct"
BN_PARTS = "bn_parts"
PIXELS = "pixls"

bn_correspondants = {
    BN_BACKGROUND: BACKGROUND,
    BN_GLOBAL: GLOBAL,
    BN_FOREGROUND: FOREGROUND,
    BN_CONCAT_PARTS: CONCAT_PARTS,
    BN_PARTS: PARTS,
}
