"""Ground-truth side of the reference's `dataset/` package (TFRecord reading); image decoding is out of scope."""
