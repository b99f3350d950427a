"""hla_mapper_b200 - ctypes mirror of libhlamap (include/hlamap.h), the B200-native hot path of
hla-mapper.  The product is the C-ABI library; this package only binds it for tests and bench.py.

    from hla_mapper_b200 import api
    db = api.Database("/path/to/hla-mapper-db")          # reference on-disk layout
    m = api.Mapper(db, device=0)                          # one context per GPU
    out = m.run(batch)                                    # screen -> trim -> score -> assign
    m.map_fastq("R1.fastq.gz", "R2.fastq.gz", "sample", "outdir")   # the reference's files
"""
