"""Entry point mirroring the reference's main.py (logging setup + the demo app), headless."""
import logging

from apps import sfm

if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s")
    sfm.main()
