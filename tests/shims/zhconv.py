"""test shim: `import zhconv` in image-ids-CTR; identity conversion"""


def convert(s, locale):
    return s
