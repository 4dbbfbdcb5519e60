"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's arithmetic for the hot path (plus the importer that
runs the real reference in the authoring container).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; the product
package (audio_fewshot_b200) never does.
"""
