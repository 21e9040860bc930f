"""`python -m ginfinity_b200 <command>` runs ginfinity_b200.cli."""
import sys

from . import cli

sys.exit(cli.main())
