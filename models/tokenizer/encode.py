"""Drop-in for the reference's models/tokenizer/encode.py (bulk encode driver, 18-47)."""
from transformer_lm_b200.encode_file import encode_file, main  # noqa: F401

if __name__ == "__main__":
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("--dataset", type=str)
    parser.add_argument("--split", type=str)
    args = parser.parse_args()
    main(args.dataset, args.split)
