import sys

from fitclip_b200.runner import main

if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
