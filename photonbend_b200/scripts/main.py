"""click group with the three commands (same names and flags as the reference's
scripts/main.py:28-35)."""

import click

from photonbend_b200.scripts.commands import alter_photo, make_pano, make_photo


@click.group()
def main():
    """Convert between fisheye photos, 360-degree double-fisheye frames and equirectangular
    panoramas on an NVIDIA B200."""


main.add_command(make_pano)
main.add_command(alter_photo)
main.add_command(make_photo)

if __name__ == "__main__":
    main()
