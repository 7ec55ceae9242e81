"""Stub of ftfy (absent in the build container) so /root/reference/clip imports.  TEST INFRASTRUCTURE."""


def fix_text(text, **kwargs):
    return text
