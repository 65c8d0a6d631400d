"""`python main.py [...]` runs the headless two-view demo (apps/sfm.py) with INFO logging, like the reference's main.py."""
import logging
import sys

from apps.sfm import main as run_demo


def _configure_logging() -> None:
    logging.basicConfig(stream=sys.stderr, level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s")


if __name__ == "__main__":
    _configure_logging()
    run_demo(sys.argv[1:])
