"""Drop-in for the reference's `tree_search.py` CLI (`python tree_search.py --config demo-config.json`):
same config keys, same growmap file written to config["dst"].  Implementation: sequoia_b200/tree_search.py."""
from sequoia_b200.tree_search import main

if __name__ == "__main__":
    main()
