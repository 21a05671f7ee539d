# Import-only stand-in so that /root/reference imports on a box without timm.
# Test infrastructure (oracle/): never imported by the product package.
